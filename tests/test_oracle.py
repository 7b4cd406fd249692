"""CPU tests of the oracle: known-answer tests, C restatement vs independent
numpy restatement, and both against the committed golden fixtures.  No GPU."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import np_oracle as no
from spectral_analyzer_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cf32(x):
    return synth.encode(np.asarray(x, np.complex128), "cf32_le")


def test_impulse_every_bin_equal():
    x = np.zeros(256, complex); x[0] = 1.0
    out = co.compute_magnitudes(cf32(x), 0, 256, "cf32_le")
    assert np.allclose(out, 20 * np.log10(1 + 1e-10), atol=1e-12)


def test_dc_lands_in_centre_bin():
    n, a = 512, 0.25
    out = co.compute_magnitudes(cf32(np.full(n, a)), 0, n, "cf32_le")
    assert abs(out[n // 2] - 20 * np.log10(a * n + 1e-10)) < 1e-9
    others = np.delete(out, n // 2)
    assert (others < -150).all()              # exact zeros give -200, rounding residue stays tiny


def test_on_bin_tone_single_bin():
    n, k = 1024, 37
    x = np.exp(2j * np.pi * k * np.arange(n) / n)
    out = co.compute_magnitudes(synth.encode(x, "cf64_le"), 0, n, "cf64_le")
    assert np.argmax(out) == (k + n // 2) % n
    assert abs(out.max() - 20 * np.log10(n)) < 1e-9


def test_integer_decode_edge_values():
    i16 = np.array([-32768, -1, 0, 1, 32767, 12345], "<i2")
    re, im = co.decode(i16, 0, 3, "ci16_le")
    assert np.array_equal(np.r_[re, im], np.array([-1.0, 0.0, 32767 / 32768, -2.0 ** -15, 2.0 ** -15, 12345 / 32768]))
    u8 = np.array([0, 127, 128, 255], np.uint8)
    re, im = co.decode(u8, 0, 2, "cu8")
    assert np.array_equal(np.r_[re, im], np.array([-127.5, 0.5, -0.5, 127.5]) / 128)
    i8 = np.array([-128, 127, -1, 0], np.int8)
    re, im = co.decode(i8, 0, 2, "ci8")
    assert np.array_equal(np.r_[re, im], np.array([-1.0, -1 / 128, 127 / 128, 0.0]))


@pytest.mark.parametrize("kind", ["ci16", "cf32", "cf64"])
def test_big_endian_twin(kind):
    x = synth.complex_signal(512, seed=5)
    le = co.compute_magnitudes(synth.encode(x, kind + "_le"), 0, 512, kind + "_le")
    be = co.compute_magnitudes(synth.encode(x, kind + "_be"), 0, 512, kind + "_be")
    assert np.array_equal(le, be)


def test_strict_reference_cf64_is_all_minus_200():
    # SpectralService.java:60-63: cf64 has no decode branch in computeMagnitudes
    raw = synth.recording(256, "cf64_le")
    out = co.compute_magnitudes(raw, 0, 256, "cf64_le", strict_reference=True)
    assert np.allclose(out, -200.0)


def test_eof_row_is_minus_150():
    raw = synth.recording(1000, "cf32_le")
    img = co.spectrogram(raw, "cf32_le", 0, 256, 256, "rect", 5)
    assert (img[3] == -150.0).all() and (img[4] == -150.0).all() and (img[2] != -150.0).any()


def test_non_power_of_two_rejected():
    with pytest.raises(RuntimeError):
        co.compute_magnitudes(synth.recording(300, "cf32_le"), 0, 300, "cf32_le")


def test_fft_matches_numpy():
    x = synth.complex_signal(4096, seed=9)
    assert np.abs(co.fft(x) - np.fft.fft(x)).max() < 1e-9


@pytest.mark.parametrize("dt", ["cf32_le", "ci16_be", "cu8", "ci8", "cf64_be"])
@pytest.mark.parametrize("win,hop", [("rect", 512), ("hann", 256), ("blackman_harris", 128)])
def test_c_vs_numpy_spectrogram(dt, win, hop):
    raw = synth.recording(512 * 9, dt, seed=2)
    a = co.spectrogram(raw, dt, 3, 512, hop, win, 20, nthreads=2)
    b = no.spectrogram(raw, dt, 3, 512, hop, win, 20)
    lin = np.abs(10 ** (a / 20) - 10 ** (b / 20))
    assert lin.max() < 1e-9 * 10 ** (b.max() / 20)


def test_golden_spectrogram():
    files = sorted(glob.glob(os.path.join(GOLD, "spec_parity_*.npz")))
    assert len(files) == 7
    for f in files:
        g = np.load(f)
        dt = str(g["datatype"])
        img = co.spectrogram(g["raw"], dt, 0, int(g["nfft"]), int(g["nfft"]), "rect", int(g["frames"]))
        assert (img[-1] == -150.0).all()
        lin = np.abs(10 ** (img / 20) - 10 ** (g["img"] / 20))
        assert lin.max() < 1e-9 * 10 ** (g["img"].max() / 20)
        # per-frame call (computeMagnitudes) equals the batched rows
        row = co.compute_magnitudes(g["raw"], 0, int(g["nfft"]), dt)
        assert np.array_equal(row, img[0])


def test_golden_c1_and_render():
    g = np.load(os.path.join(GOLD, "spec_c1_mini.npz"))
    img = co.spectrogram(g["raw"], "cf32_le", 0, 1024, 512, "hann", 9)
    assert np.abs(img - g["img"]).max() < 1e-6
    r = np.load(os.path.join(GOLD, "render_c1_mini.npz"))
    for name in ("Heatmap", "Grayscale"):
        rgba = co.render_rgba(g["img"], float(r["fs"]), -160.0, -30.0, name)
        assert np.array_equal(rgba, r[name.lower()])
        assert np.array_equal(rgba, no.render_rgba(g["img"], float(r["fs"]), -160.0, -30.0, name))


def test_render_canvas_row_pick_and_flip():
    db = np.tile(np.linspace(-60, 60, 64), (3, 1))
    full = co.render_rgba(db, 1e6, -160.0, -30.0, "Grayscale")
    can = co.render_canvas(db, 16, 1e6, -160.0, -30.0, "Grayscale")
    for f in range(16):
        assert np.array_equal(can[16 - 1 - f, :, :], full[:, int(f / 16 * 64), :])


def test_heatmap_breakpoints():
    conv = 10 * np.log10(1.0 / 64) + 20 * np.log10(64)
    mk = lambda n: np.full((1, 64), -160.0 + 130.0 * n + conv)
    px = lambda n: co.render_rgba(mk(n), 1.0, -160.0, -30.0, "Heatmap")[0, 0]
    assert tuple(px(0.1)) == (0, 0, 0, 255)
    assert tuple(px(0.35)) == (128, 0, 128, 255) or tuple(px(0.35)) == (127, 0, 128, 255)
    assert tuple(px(0.75)) == (255, 128, 0, 255) or tuple(px(0.75)) == (255, 127, 0, 255)
    assert tuple(px(1.5)) == (255, 255, 0, 255)


def test_downconvert_and_welch_golden():
    g = np.load(os.path.join(GOLD, "analysis_mini.npz"))
    dc = co.downconvert(g["raw"], "cf32_le", 100, 36000, 0.125, 4, False)
    dcf = co.downconvert(g["raw"], "cf32_le", 100, 36000, 0.125, 4, True)
    assert dc.shape == (2, 9000)
    assert np.abs(dc - g["dc"]).max() < 1e-12 and np.abs(dcf - g["dcf"]).max() < 1e-12
    psd = co.psd_welch(dc, 1e6 / 4, 2048)
    assert np.abs(psd - g["psd"]).max() < 1e-8
    # the 0.125 cycles/sample tone was moved to DC: PSD peak at 0 Hz
    assert abs(psd[0][np.argmax(psd[1])]) < 1e6 / 4 / 2048 * 1.5


def test_lowpass_taps():
    h = co.lowpass_taps(16)
    assert h.size == 129 and abs(h.sum() - 1) < 1e-12 and np.allclose(h, h[::-1])
    assert np.abs(h - no.lowpass_taps(16)).max() < 1e-15
    H = np.abs(np.fft.fft(h, 8192))
    assert 20 * np.log10(H[int(8192 * 1.5 / 16)]) < -40          # stop band


def test_downconvert_tone_to_dc():
    n = 1 << 15
    x = 0.5 * np.exp(2j * np.pi * 0.2 * np.arange(n))
    out = co.downconvert(synth.encode(x, "cf64_le"), "cf64_le", 0, n, 0.2, 16, False)
    z = out[0] + 1j * out[1]
    assert np.abs(z[20:] - 0.5).max() < 1e-9


# ---- rows next to the hot path: IqData packers, analysis series ----
def test_iq_pack_known_answers_and_golden():
    """IqData.getInterleavedBinary (IqData.java:160-187): Java narrowing (short)(32767*x)."""
    iq = np.array([[1.0, -1.0, 0.5, 1.00002, 3.2, 1e12, np.nan], [0.0, 2.0 ** -15, -0.5, -1.00004, -7.9, -1e12, -0.0]])
    i16 = np.frombuffer(co.iq_pack(iq, "int16"), "<i2").reshape(-1, 2)
    # 32767*1.00002 = 32767.65 -> 32767 ; 32767*3.2 = 104854.4 -> low 16 bits of 104854 = -26218 (wraps, no clipping);
    # 1e12 saturates the int conversion (0x7fffffff -> low 16 bits = -1); NaN -> 0
    assert i16[:, 0].tolist() == [32767, -32767, 16383, 32767, 104854 - 131072, -1, 0]
    assert i16[:, 1].tolist() == [0, 0, -16383, -32768, -258859 + 262144, 0, 0]
    f32 = np.frombuffer(co.iq_pack(iq, "float32"), "<f4").reshape(-1, 2)
    assert np.array_equal(f32[:5, 0], iq[0, :5].astype(np.float32)) and np.isnan(f32[6, 0])
    g = np.load(os.path.join(GOLD, "iqdata_mini.npz"))
    for fmt, key in (("float32", "edge_f32"), ("int16", "edge_i16")):
        assert co.iq_pack(g["edge"], fmt) == no.iq_pack(g["edge"], fmt) == g[key].tobytes()


def test_analysis_series_known_answers_and_golden():
    """updateMagnitudeChart / updateFrequencyChart (AnalysisDialogController.java:219-290)."""
    n, fs, f0 = 4000, 1.0e6, 12.5e3
    z = 0.25 * np.exp(2j * np.pi * f0 / fs * np.arange(n))
    mag, frq = co.analysis_series(np.stack([z.real, z.imag]), fs, 0.3, 0.1, 100e6)
    assert np.abs(mag - 20 * np.log10(0.25)).max() < 1e-9            # constant envelope: the EMA is the envelope
    assert np.isnan(frq[0]) and np.abs(frq[1:] - (100e6 + f0)).max() < 1e-3
    # phase wrap: a tone at -0.45 fs steps the phase by -0.9 pi; +0.55 fs would alias to the same value
    z = np.exp(-2j * np.pi * 0.45 * np.arange(64))
    _, frq = co.analysis_series(np.stack([z.real, z.imag]), 2.0, 1.0, 1.0, 0.0)
    assert np.abs(frq[1:] + 0.9).max() < 1e-9
    # alpha = 1 -> no smoothing; first value initialises the average (:234-235, :276-277)
    x = np.array([[3.0, 0.0, 0.0], [4.0, 1.0, 0.0]])
    mag, _ = co.analysis_series(x, 1.0, 0.5, 1.0, 0.0)
    assert np.allclose(mag[:2], 20 * np.log10([5.0, 3.0])) and np.isclose(mag[2], 20 * np.log10(1.5))
    g = np.load(os.path.join(GOLD, "iqdata_mini.npz"))
    a = np.load(os.path.join(GOLD, "analysis_mini.npz"))
    cm, cf = co.analysis_series(a["dc"], 250e3, 0.2, 0.05, 915e6)
    nm, nf = no.analysis_series(a["dc"], 250e3, 0.2, 0.05, 915e6)
    assert np.abs(cm - g["mag"]).max() < 1e-10 and np.abs(cf[1:] - g["frq"][1:]).max() < 1e-5
    assert np.abs(cm - nm).max() < 1e-10 and np.abs(cf[1:] - nf[1:]).max() < 1e-5


@pytest.mark.parametrize("n", [63, 777, 1000, 1024])
@pytest.mark.parametrize("scaling,detrend", [("density", None), ("spectrum", "constant"), ("density", "constant")])
def test_welch_any_length_against_scipy(n, scaling, detrend):
    """ora_psd_welch_ex (any transform length; the reference's short-signal branch passes nfft = N,
    AnalysisDialogController.java:304-307) against scipy.signal.welch: segments, scaling, mean removal, axis."""
    import scipy.signal as ss
    rng = np.random.default_rng(n)
    x = rng.normal(size=3 * n + 5) + 1j * rng.normal(size=3 * n + 5) + (1.5 - 0.5j)
    iq = np.stack([x.real, x.imag])
    hop = max(1, n // 4)
    f, d = co.psd_welch(iq, 1e3, n, hop=hop, window="hann", cfg=co.analysis_cfg(scaling=scaling, detrend=detrend))
    fr, p = ss.welch(x, fs=1e3, window=ss.get_window("hann", n, fftbins=True), nperseg=n, noverlap=n - hop, nfft=n,
                     detrend=detrend if detrend else False, return_onesided=False, scaling=scaling)
    assert np.abs(d - 10 * np.log10(np.fft.fftshift(p) + 1e-30)).max() < 1e-8
    assert np.allclose(f, np.fft.fftshift(fr))


@pytest.mark.parametrize("delay", ["causal", "same", "valid"])
@pytest.mark.parametrize("length", ["floor", "ceil"])
@pytest.mark.parametrize("down,ntaps", [(4, None), (4, 21), (7, 100)])
def test_downconvert_profile_against_scipy(delay, length, down, ntaps):
    """ora_downconvert_ex: taps, delay compensation and length rule against numpy / scipy.signal.lfilter."""
    import scipy.signal as ss
    count, start, f = 5003, 11, 0.0917
    raw = synth.recording(count + start, "cf32_le", seed=12)
    taps = None
    if ntaps:
        taps = ss.firwin(ntaps, 1.0 / down)
    cfg = co.analysis_cfg(taps=taps, delay=delay, length=length)
    got = co.downconvert_ex(raw, "cf32_le", start, count, f, down, False, cfg)
    h = taps if taps is not None else co.lowpass_taps(down)
    x = np.frombuffer(raw.tobytes(), np.complex64)[start:start + count].astype(np.complex128)
    y = x * np.exp(-2j * np.pi * np.mod(f * np.arange(count), 1.0))
    L = len(h)
    off = {"causal": 0, "same": (L - 1) // 2, "valid": L - 1}[delay]
    full = ss.lfilter(h, 1.0, np.concatenate([y, np.zeros(off + down)]))       # full[n] = sum h[k] y[n - k], zero tail
    if delay == "valid":
        m = (count - L) // down + 1 if count >= L else 0
    else:
        m = count // down if length == "floor" else -(-count // down)
    ref = full[off + down * np.arange(m)]
    assert got.shape == (2, m)
    assert np.abs(got[0] + 1j * got[1] - ref).max() < 1e-12
