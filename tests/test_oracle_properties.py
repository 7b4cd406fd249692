"""Randomised agreement of the two independent restatements (oracle/sa_oracle.c, plain C, and oracle/np_oracle.py,
numpy / scipy) and the properties the domain offers, on CPU: the checker is checked before it checks the kernels."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as co
from oracle import np_oracle as no
from spectral_analyzer_b200 import sigmf

DTYPES = ["cf32_le", "cf32_be", "ci16_le", "ci16_be", "cu8", "ci8", "cf64_le", "cf64_be"]
WINDOWS = ["rect", "hann", "hamming", "blackman", "blackman_harris"]


def random_recording(rng, n, dt):
    bps = sigmf.bytes_per_sample(dt)
    if dt.startswith("cf32"):
        a = rng.standard_normal(2 * n).astype(">f4" if dt.endswith("_be") else "<f4")
    elif dt.startswith("cf64"):
        a = rng.standard_normal(2 * n).astype(">f8" if dt.endswith("_be") else "<f8")
    elif dt.startswith("ci16"):
        a = rng.integers(-32768, 32768, 2 * n).astype(">i2" if dt.endswith("_be") else "<i2")
    else:
        a = rng.integers(0, 256, 2 * n).astype(np.uint8)
    raw = np.frombuffer(a.tobytes(), np.uint8)
    assert raw.size == n * bps
    return raw


@settings(max_examples=60, deadline=None)
@given(dt=st.sampled_from(DTYPES), win=st.sampled_from(WINDOWS), log2n=st.integers(6, 11), hop=st.integers(1, 3000),
       start=st.integers(0, 50), frames=st.integers(1, 6), seed=st.integers(0, 1 << 30), mode=st.sampled_from([0, 1]))
def test_c_and_numpy_spectrogram_agree(dt, win, log2n, hop, start, frames, seed, mode):
    nfft = 1 << log2n
    rng = np.random.default_rng(seed)
    n = start + (frames - 1) * hop + nfft - (nfft // 3 if seed % 4 == 0 else 0)      # sometimes the last frame is past EOF
    raw = random_recording(rng, max(n, 1), dt)
    a = co.spectrogram(raw, dt, start, nfft, hop, win, frames, db_mode=mode)
    b = no.spectrogram(raw, dt, start, nfft, hop, win, frames, db_mode=mode)
    assert a.shape == b.shape == (frames, nfft)
    eof = a[:, 0] == -150.0
    assert np.array_equal(eof, b[:, 0] == -150.0) and (a[eof] == -150.0).all()
    # two FP64 FFTs with different operation orders: compare in linear amplitude against the frame's maximum
    la, lb = 10.0 ** (a[~eof] / 20.0), 10.0 ** (b[~eof] / 20.0)
    if la.size:
        assert np.abs(la - lb).max() <= 1e-12 * max(la.max(), 1e-10) + 1e-22


@settings(max_examples=40, deadline=None)
@given(dt=st.sampled_from(["cf32_le", "ci16_be", "cu8", "ci8", "cf64_le"]), down=st.integers(1, 40),
       count=st.integers(0, 5000), start=st.integers(0, 100), f=st.floats(-0.5, 0.5), fast=st.booleans(),
       seed=st.integers(0, 1 << 30))
def test_c_and_numpy_downconvert_agree(dt, down, count, start, f, fast, seed):
    rng = np.random.default_rng(seed)
    raw = random_recording(rng, start + count + 3, dt)
    a = co.downconvert(raw, dt, start, count, f, down, fast)
    b = no.downconvert(raw, dt, start, count, f, down, fast)
    assert a.shape == b.shape == (2, count // down)
    if a.size:
        scale = max(np.abs(b).max(), 1e-3)
        assert np.abs(a - b).max() <= 1e-9 * scale


@settings(max_examples=40, deadline=None)
@given(log2n=st.integers(6, 10), seed=st.integers(0, 1 << 30), k=st.integers(0, 63), amp=st.floats(1e-3, 10.0))
def test_spectrogram_properties(log2n, seed, k, amp):
    """Parseval, on-bin tone, and shift by one hop = next row (frame indexing is exact)."""
    nfft = 1 << log2n
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(3 * nfft) + 1j * rng.standard_normal(3 * nfft)) * amp
    iq = np.empty(6 * nfft, np.float64)
    iq[0::2], iq[1::2] = x.real, x.imag
    raw = np.frombuffer(iq.tobytes(), np.uint8)
    rows = co.spectrogram(raw, "cf64_le", 0, nfft, nfft // 2, "rect", 5, db_mode=1)       # 10 log10(|X|^2 + 1e-20)
    p = 10.0 ** (rows / 10.0) - 1e-20
    for t in range(5):
        seg = x[t * nfft // 2: t * nfft // 2 + nfft]
        assert abs(p[t].sum() / (nfft * (np.abs(seg) ** 2).sum()) - 1) < 1e-9
    shifted = co.spectrogram(raw, "cf64_le", nfft // 2, nfft, nfft // 2, "rect", 4, db_mode=1)
    assert np.array_equal(shifted, rows[1:])
    tone = amp * np.exp(2j * np.pi * k * np.arange(nfft) / nfft)
    tq = np.empty(2 * nfft, np.float64)
    tq[0::2], tq[1::2] = tone.real, tone.imag
    r = co.spectrogram(np.frombuffer(tq.tobytes(), np.uint8), "cf64_le", 0, nfft, nfft, "rect", 1)[0]
    assert int(np.argmax(r)) == (k + nfft // 2) % nfft                                   # fft-shifted, SpectralService.java:78
    assert abs(r.max() - 20 * np.log10(amp * nfft + 1e-10)) < 1e-9


@settings(max_examples=40, deadline=None)
@given(log2n=st.integers(6, 10), extra=st.integers(0, 3000), hopdiv=st.sampled_from([1, 2, 4, 8]),
       win=st.sampled_from(WINDOWS), seed=st.integers(0, 1 << 30), fs=st.floats(1.0, 1e8))
def test_c_and_numpy_welch_agree(log2n, extra, hopdiv, win, seed, fs):
    nfft = 1 << log2n
    rng = np.random.default_rng(seed)
    iq = rng.standard_normal((2, nfft + extra))
    a = co.psd_welch(iq, fs, nfft, hop=nfft // hopdiv, window=win)
    b = no.psd_welch(iq, fs, nfft, hop=nfft // hopdiv, win=win)
    assert np.allclose(a[0], b[0], rtol=1e-12, atol=1e-9 * fs / nfft)            # frequency axis centred on 0
    assert np.abs(a[1] - b[1]).max() < 1e-9                                      # dB/Hz


@settings(max_examples=60, deadline=None)
@given(n=st.integers(0, 3000), seed=st.integers(0, 1 << 30), scale=st.sampled_from([0.1, 1.0, 1.5, 40.0, 1e6]),
       fmt=st.sampled_from(["float32", "int16"]))
def test_c_and_numpy_iq_pack_agree_bit_for_bit(n, seed, scale, fmt):
    rng = np.random.default_rng(seed)
    iq = rng.standard_normal((2, n)) * scale
    if n > 3:
        iq[0, 0], iq[1, 1], iq[0, 2] = 1.0, -1.0, 32767.5 / 32767.0             # edges of the (short)(32767 x) narrowing
    assert co.iq_pack(iq, fmt) == no.iq_pack(iq, fmt)


@settings(max_examples=40, deadline=None)
@given(n=st.integers(2, 4000), seed=st.integers(0, 1 << 30), am=st.floats(0.01, 1.0), af=st.floats(0.01, 1.0),
       fs=st.floats(1e3, 1e8), fc=st.floats(-1e9, 1e9))
def test_c_and_numpy_analysis_series_agree(n, seed, am, af, fs, fc):
    rng = np.random.default_rng(seed)
    iq = rng.standard_normal((2, n)) + 0.1
    gm, gf = co.analysis_series(iq, fs, am, af, fc)
    rm, rf = no.analysis_series(iq, fs, am, af, fc)
    assert np.abs(gm - rm).max() < 1e-9
    assert np.isnan(gf[0]) and np.isnan(rf[0])
    assert np.abs(gf[1:] - rf[1:]).max() <= 1e-9 * fs + 1e-6 * abs(fc) * 1e-9 + 1e-6
