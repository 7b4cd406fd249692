"""Seeded random sweep of the spectrogram entry point against the oracle: every nfft of the slider range, every
decode branch, random hops (overlapping, gapped, odd), start offsets, windows, both dB modes, frames that run
past EOF.  Kernel selection depends on alignment (base, start and hop bytes modulo 16), so odd parameters and
aligned ones land on different kernels; all must agree with the same restatement."""
import numpy as np
import pytest

from oracle import c_oracle as co
from spectral_analyzer_b200 import synth
from util import check_db_parity

pytestmark = pytest.mark.gpu
SIZES = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
DTYPES = ["cf32_le", "cf32_be", "ci16_le", "ci16_be", "cu8", "ci8"]
WINDOWS = ["rect", "hann", "hamming", "blackman", "blackman_harris"]


def cases():
    rng = np.random.default_rng(20261018)
    out = []
    for i in range(66):
        nfft = SIZES[i % len(SIZES)]
        dt = DTYPES[int(rng.integers(len(DTYPES)))]
        win = WINDOWS[int(rng.integers(len(WINDOWS)))]
        kind = int(rng.integers(4))
        hop = [nfft, nfft // 2, int(rng.integers(1, 2 * nfft)), nfft // 4 + 1][kind]
        start = int(rng.integers(0, 50)) if rng.random() < 0.7 else 0
        frames = int(rng.integers(2, 9 if nfft <= 8192 else 5))
        short = int(rng.integers(0, nfft)) if rng.random() < 0.4 else 0          # cut the recording: EOF rows
        mode = int(rng.random() < 0.25)
        out.append((nfft, dt, win, hop, start, frames, short, mode))
    return out


@pytest.mark.parametrize("nfft,dt,win,hop,start,frames,short,mode", cases())
def test_random_parameters_match_oracle(engine, nfft, dt, win, hop, start, frames, short, mode):
    n = max(start + (frames - 1) * hop + nfft - short, 1)
    raw = synth.recording(n, dt, seed=nfft + hop)
    ref = co.spectrogram(raw, dt, start, nfft, hop, win, frames, db_mode=mode)
    got = engine.spectrogram(raw, dt, nfft, frames, hop=hop, window=win, start_sample=start, db_mode=mode)
    eof = np.array([start + t * hop + nfft > n for t in range(frames)])
    assert (got[eof] == -150.0).all() and (ref[eof] == -150.0).all()
    if (~eof).any():
        check_db_parity(got[~eof], ref[~eof], mode_power=bool(mode))


def dc_cases():
    rng = np.random.default_rng(77)
    out = []
    for i in range(40):
        down = int([rng.integers(1, 9), rng.integers(9, 65), rng.integers(65, 300), rng.integers(300, 900)][i % 4])
        dt = DTYPES[int(rng.integers(len(DTYPES)))] if i % 5 else "cf64_le"
        fast = bool(rng.random() < 0.3)
        count = int(rng.integers(20 * down, 60 * down + 5000))
        start = int(rng.integers(0, 3000))
        f = float(rng.uniform(-0.5, 0.5)) if i % 7 else float(rng.integers(-3, 4))          # integer offsets wrap to 0
        out.append((dt, down, fast, start, count, f))
    return out


@pytest.mark.parametrize("dt,down,fast,start,count,f", dc_cases())
def test_random_downconvert_matches_oracle(engine, dt, down, fast, start, count, f):
    """Batches of mixed decimation factors exercise the launch grouping (taps in the parameter bank for D <= 256,
    from global memory above, warp-per-output kernel for D > 512) in one call."""
    raw = synth.recording(start + count + 17, dt, seed=down)
    ref = co.downconvert(raw, dt, start, count, f, down, fast)
    got = engine.downconvert(raw, dt, start, count, f, down, fast)
    assert got.shape == ref.shape
    if ref.size:
        # 1e-5 of full scale (the recordings peak near 0.9): a strongly attenuated output is still held to the
        # absolute error of FP32 taps, NCO and accumulation over up to 8*down+1 products
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 0.5)


def test_random_batch_mixed_decimations(engine):
    rng = np.random.default_rng(5)
    raw = synth.recording(400_000, "ci16_le", seed=9)
    anns = []
    for i in range(37):
        down = int(rng.choice([1, 3, 16, 16, 40, 257, 600]))
        count = int(rng.integers(30 * down, 30 * down + 40_000))
        anns.append((int(rng.integers(0, 400_000 - count)), count, float(rng.uniform(-0.4, 0.4)), down, bool(i % 6 == 0)))
    iq, psd = engine.downconvert_psd_batch(raw, "ci16_le", 2.0e6, anns, psd_nfft=256)
    for (s, c, f, d, fast), z, p in zip(anns, iq, psd):
        ref = co.downconvert(raw, "ci16_le", s, c, f, d, fast)
        assert np.abs(z - ref).max() <= 1e-5 * max(np.abs(ref).max(), 0.5)
        nf = min(256, ref.shape[1])                 # the caller's short-signal rule: one window of M points, rest of the row NaN
        if nf:
            rp = co.psd_welch(ref, 2.0e6 / d, nf)[1]
            top = rp > rp.max() - 40
            assert np.abs(p[:nf] - rp)[top].max() < 5e-3
        assert np.isnan(p[nf:]).all()


def profile_cases():
    """Random analysis profiles: tap counts around the 8*down+1 boundary between the staged and the warp-per-output
    kernel, every delay x length rule, spans shorter than the filter, starts at 0 and ends at the end of the buffer."""
    rng = np.random.default_rng(4242)
    out = []
    for i in range(48):
        down = int([rng.integers(1, 9), rng.integers(9, 33), rng.integers(33, 200), rng.integers(200, 700)][i % 4])
        ntaps = [None, int(rng.integers(1, 8 * down + 2)), 8 * down + 1, int(rng.integers(8 * down + 2, 8 * down + 300))][int(rng.integers(4))]
        delay = ["causal", "same", "valid"][int(rng.integers(3))]
        length = ["floor", "ceil"][int(rng.integers(2))]
        dt = DTYPES[int(rng.integers(len(DTYPES)))] if i % 6 else "cf64_le"
        count = int(rng.integers(1, 40 * down + 3000)) if i % 9 else int(rng.integers(1, 3 * down + 5))
        start = 0 if i % 4 == 0 else int(rng.integers(0, 2000))
        out.append((dt, down, ntaps, delay, length, start, count, float(rng.uniform(-0.5, 0.5))))
    return out


@pytest.mark.parametrize("dt,down,ntaps,delay,length,start,count,f", profile_cases())
def test_random_analysis_profiles_match_oracle(engine, dt, down, ntaps, delay, length, start, count, f):
    raw = synth.recording(start + count, dt, seed=down + count)                 # the annotation ends at the end of the buffer
    taps = None
    if ntaps:
        taps = np.random.default_rng(ntaps).standard_normal(ntaps)
        taps /= np.abs(taps).sum()
    cfg = co.analysis_cfg(taps=taps, delay=delay, length=length)
    engine.set_analysis_config(taps=taps, delay=delay, length=length)
    try:
        for fast in (False, True):
            ref = co.downconvert_ex(raw, dt, start, count, f, down, fast, cfg)
            got = engine.downconvert(raw, dt, start, count, f, down, fast)
            assert got.shape == ref.shape, (fast, got.shape, ref.shape)
            if ref.size:
                assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 0.5), fast
    finally:
        engine.reset_analysis_config()


def psd_cases():
    rng = np.random.default_rng(99)
    out = []
    for i in range(24):
        nfft = int([rng.integers(1, 64), rng.integers(64, 1200), 2 ** int(rng.integers(6, 14)), rng.integers(1200, 9000)][i % 4])
        segs = int(rng.integers(1, 6))
        hop = max(1, int(nfft * rng.uniform(0.1, 1.2)))
        n = nfft + (segs - 1) * hop + int(rng.integers(0, hop))
        out.append((nfft, hop, n, WINDOWS[int(rng.integers(len(WINDOWS)))], ["density", "spectrum"][int(rng.integers(2))],
                    [None, "constant"][int(rng.integers(2))], ["f32", "f64"][i % 2]))
    return out


@pytest.mark.parametrize("nfft,hop,n,window,scaling,detrend,prec", psd_cases())
def test_random_psd_requests_match_oracle(engine, nfft, hop, n, window, scaling, detrend, prec):
    rng = np.random.default_rng(nfft + n)
    x = (0.4 - 0.3j) + 0.5 * np.exp(2j * np.pi * 0.123 * np.arange(n)) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.stack([x.real, x.imag])
    ref = co.psd_welch(iq, 48e3, nfft, hop=hop, window=window, cfg=co.analysis_cfg(scaling=scaling, detrend=detrend))
    engine.set_analysis_config(psd_scaling=scaling, psd_detrend=detrend, psd_precision=prec)
    try:
        got = engine.psd_welch(iq, 48e3, nfft, hop=hop, window=window)
    finally:
        engine.reset_analysis_config()
    assert got.shape == (2, nfft) and np.allclose(got[0], ref[0], rtol=0, atol=1e-9)
    ok = np.isfinite(ref[1])
    assert np.array_equal(np.isfinite(got[1]), ok)
    if ok.any():
        top = ok & (ref[1] > np.nanmax(ref[1][ok]) - (60 if prec == "f64" else 40))
        tol = 1e-8 if prec == "f64" else 2e-3
        assert np.abs(got[1] - ref[1])[top].max() < tol
