"""GPU parity tests of the annotation-analysis path (downconvert + Welch PSD) against the oracle's
self-defined spec (JDSP is not vendored in the reference: parity unpinned)."""
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from spectral_analyzer_b200 import (synth, ExtractDownConvertService, AsyncExtractDownConvertService,
                                    PowerSpectralDensity, EngineError)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DC_TOL = 1e-5          # relative to the largest output magnitude (FP32 taps, accumulation and NCO)
PSD_TOL_DB = 2e-3      # dB on bins within 60 dB of the PSD peak


def rel_err(got, ref):
    return np.abs(got - ref).max() / np.abs(ref).max()


@pytest.mark.parametrize("dt", ["cf32_le", "cf32_be", "ci16_le", "cu8", "ci8", "cf64_le"])
@pytest.mark.parametrize("down,fast", [(16, False), (16, True), (4, False), (1, False), (7, False), (7, True),
                                       (100, False), (600, False), (3000, True)])
def test_downconvert_matches_oracle(engine, dt, down, fast):
    count = max(20000, 40 * down)
    raw = synth.recording(count + 300, dt, seed=4)
    ref = co.downconvert(raw, dt, 123, count, -0.28137, down, fast)
    got = engine.downconvert(raw, dt, 123, count, -0.28137, down, fast)
    assert got.shape == ref.shape == (2, count // down)
    assert rel_err(got, ref) < DC_TOL


def test_downconvert_service_and_async(engine):
    svc = ExtractDownConvertService(engine)
    raw = synth.recording(1 << 16, "ci16_le", seed=6)
    a = svc.extractAndDownConvert(raw, 1000, 50000, "ci16_le", 0.125, 16)          # default = conventional
    b = svc.extractAndDownConvert(raw, 1000, 50000, "ci16_le", 0.125, 16, False)
    assert np.array_equal(a, b) and a.shape == (2, 3125)
    asvc = AsyncExtractDownConvertService(engine, workers=4)
    futs = [asvc.extractAndDownConvertAsync(raw, 1000 + 10 * i, 20000, "ci16_le", 0.125, 16, bool(i & 1)) for i in range(8)]
    for i, f in enumerate(futs):
        ref = co.downconvert(raw, "ci16_le", 1000 + 10 * i, 20000, 0.125, 16, bool(i & 1))
        assert rel_err(f.result(), ref) < DC_TOL
    with pytest.raises(EngineError) as ei:
        svc.extractAndDownConvert(raw, 60000, 50000, "ci16_le", 0.1, 16)            # IndexOutOfBounds in Java
    assert ei.value.code == 3


def test_long_nco_phase_is_exact(engine):
    """2^24 samples: an FP32 phase recurrence would drift; the 64-bit accumulator must not."""
    n = 1 << 24
    f = 0.2001
    x = 0.5 * np.exp(2j * np.pi * np.mod(f * np.arange(n), 1.0))
    got = engine.downconvert(synth.encode(x, "cf32_le"), "cf32_le", 0, n, f, 256, False)
    z = got[0] + 1j * got[1]
    assert np.abs(z[16:] - 0.5).max() < 5e-6
    assert np.abs(z[-1000:] - 0.5).max() < 5e-6


@pytest.mark.parametrize("nfft", [64, 1024, 8192, 16384])
def test_psd_welch_matches_oracle(engine, nfft):
    rng = np.random.default_rng(nfft)
    n = nfft * 6 + 37
    t = np.arange(n)
    x = 0.3 * np.exp(2j * np.pi * 0.11 * t) + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.stack([x.real, x.imag])
    ref = co.psd_welch(iq, 2.5e5, nfft)
    PowerSpectralDensity.engine = engine
    got = PowerSpectralDensity.calculatePsdWelch(iq, 2.5e5, nfft)
    assert got.shape == (2, nfft)
    assert np.allclose(got[0], ref[0], rtol=0, atol=1e-9)
    strong = ref[1] > ref[1].max() - 60
    assert np.abs(got[1] - ref[1])[strong].max() < PSD_TOL_DB
    assert abs(got[0][np.argmax(got[1])] - 0.11 * 2.5e5) <= 2.5e5 / nfft


def test_psd_rejects_bad_sizes(engine):
    iq = np.zeros((2, 1000))
    with pytest.raises(EngineError):
        engine.psd_welch(iq, 1.0, 2048)          # longer than the signal
    with pytest.raises(EngineError):
        engine.psd_welch(iq, 1.0, 0)
    with pytest.raises(EngineError):
        engine.psd_welch(np.zeros((2, 40000)), 1.0, 40000)      # beyond the direct kernel's shared memory


@pytest.mark.parametrize("n", [1, 2, 63, 777, 1000, 4097, 8191])
@pytest.mark.parametrize("prec,tol", [("f32", 1e-3), ("f64", 1e-9)])
def test_psd_of_a_short_signal_any_length(engine, n, prec, tol):
    """The reference's short-signal branch (AnalysisDialogController.java:304-307): psdNfft = N for N < 8192, one
    window over the whole signal, any N.  Direct DFT kernel against the checker (which agrees with scipy.signal.welch
    for non powers of two, tests/test_oracle.py)."""
    rng = np.random.default_rng(n)
    t = np.arange(n)
    x = 0.3 * np.exp(2j * np.pi * 0.11 * t) + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.stack([x.real, x.imag])
    win = "hann" if n > 2 else "rect"            # the periodic Hann window of 1 or 2 points has no energy
    ref = co.psd_welch(iq, 2.5e5, n, window=win)
    engine.set_analysis_config(psd_precision=prec)
    try:
        got = engine.psd_welch(iq, 2.5e5, n, window=win)
    finally:
        engine.reset_analysis_config()
    assert got.shape == (2, n) and np.allclose(got[0], ref[0], rtol=0, atol=1e-9)
    strong = ref[1] > ref[1].max() - (60 if prec == "f64" else 40)       # FP32 sums of up to 8191 products
    assert np.abs(got[1] - ref[1])[strong].max() < tol


@pytest.mark.parametrize("nfft,hop", [(1000, 250), (384, 100), (1536, 1536)])
@pytest.mark.parametrize("scaling,detrend", [("density", None), ("spectrum", "constant"), ("density", "constant")])
def test_psd_any_length_segments_scaling_detrend(engine, nfft, hop, scaling, detrend):
    """Several segments of a non power-of-two length, both scalings, mean removal: engine == checker == scipy."""
    import scipy.signal as ss
    rng = np.random.default_rng(nfft)
    n = nfft * 5 + 11
    x = (0.7 + 0.2j) + 0.3 * np.exp(2j * np.pi * 0.2 * np.arange(n)) + 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.stack([x.real, x.imag])
    ref = co.psd_welch(iq, 1e4, nfft, hop=hop, window="hamming", cfg=co.analysis_cfg(scaling=scaling, detrend=detrend))
    _, p = ss.welch(x, fs=1e4, window=ss.get_window("hamming", nfft, fftbins=True), nperseg=nfft, noverlap=nfft - hop, nfft=nfft,
                    detrend=detrend if detrend else False, return_onesided=False, scaling=scaling)
    assert np.abs(ref[1] - 10 * np.log10(np.fft.fftshift(p) + 1e-30)).max() < 1e-8
    for prec, tol in (("f32", 1e-3), ("f64", 1e-9)):
        engine.set_analysis_config(psd_scaling=scaling, psd_detrend=detrend, psd_precision=prec)
        try:
            got = engine.psd_welch(iq, 1e4, nfft, hop=hop, window="hamming")
        finally:
            engine.reset_analysis_config()
        strong = ref[1] > ref[1].max() - (60 if prec == "f64" else 40)
        assert np.abs(got[1] - ref[1])[strong].max() < tol, (prec, np.abs(got[1] - ref[1])[strong].max())


@pytest.mark.parametrize("nfft", [64, 256, 1024, 4096, 8192])
@pytest.mark.parametrize("scaling,detrend", [("density", None), ("spectrum", "constant")])
def test_psd_welch_fp64_and_knobs_on_the_stockham_kernels(engine, nfft, scaling, detrend):
    """Power-of-two lengths: FP64 transforms meet 1e-9 dB on the top 60 dB (north_star: 'tighter for the FP64 path'),
    FP32 transforms 1e-3 dB on the top 40 dB, for both scalings, with and without mean removal."""
    rng = np.random.default_rng(nfft + 1)
    n = nfft * 7 + 5
    x = (0.25 - 0.1j) + 0.3 * np.exp(2j * np.pi * 0.11 * np.arange(n)) + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.stack([x.real, x.imag])
    ref = co.psd_welch(iq, 2.5e5, nfft, cfg=co.analysis_cfg(scaling=scaling, detrend=detrend))
    for prec, tol, span in (("f64", 1e-9, 60), ("f32", 1e-3, 40)):
        engine.set_analysis_config(psd_scaling=scaling, psd_detrend=detrend, psd_precision=prec)
        try:
            got = engine.psd_welch(iq, 2.5e5, nfft)
        finally:
            engine.reset_analysis_config()
        strong = ref[1] > ref[1].max() - span
        assert np.abs(got[1] - ref[1])[strong].max() < tol, (prec, np.abs(got[1] - ref[1])[strong].max())


@pytest.mark.parametrize("dt", ["cf32_le", "ci16_be", "cu8", "cf64_le"])
@pytest.mark.parametrize("delay,length", [("causal", "ceil"), ("same", "floor"), ("same", "ceil"), ("valid", "floor")])
@pytest.mark.parametrize("down,ntaps", [(16, None), (16, 65), (16, 129), (5, 33), (4, 101), (700, None)])
def test_downconvert_profile_taps_delay_length(engine, dt, delay, length, down, ntaps):
    """The JDSP unknowns as parameters: caller taps (shorter than, equal to and longer than 8*down+1: the staged,
    pipelined and warp-per-output kernels), group-delay compensation, output-length rule.  count is not a multiple
    of down, and the annotation ends at the end of the buffer (zero tail, no read past it)."""
    count = max(20011, 40 * down + 3)
    raw = synth.recording(count + 77, dt, seed=9)
    taps = None
    if ntaps:
        k = np.arange(ntaps) - (ntaps - 1) / 2
        taps = np.sinc(k / down) * np.blackman(ntaps)
        taps /= taps.sum()
    cfg = co.analysis_cfg(taps=taps, delay=delay, length=length)
    ref = co.downconvert_ex(raw, dt, 77, count, 0.1713, down, False, cfg)
    engine.set_analysis_config(taps=taps, delay=delay, length=length)
    try:
        assert engine.downconvert_length(count, down) == ref.shape[1]
        got = engine.downconvert(raw, dt, 77, count, 0.1713, down, False)
        fast = engine.downconvert(raw, dt, 77, count, 0.1713, down, True)
    finally:
        engine.reset_analysis_config()
    assert got.shape == ref.shape
    # 1e-5 of full scale (the tones are far from the NCO frequency here: once the start-up transient is excluded, as in
    # the valid mode, the outputs are ~2e-3 and still held to the absolute error of FP32 taps, NCO and accumulation)
    assert np.abs(got - ref).max() <= DC_TOL * max(np.abs(ref).max(), 0.5)
    ref_fast = co.downconvert_ex(raw, dt, 77, count, 0.1713, down, True, cfg)
    assert fast.shape == ref_fast.shape and np.abs(fast - ref_fast).max() <= DC_TOL * max(np.abs(ref_fast).max(), 0.5)


def test_analysis_config_round_trip_and_validation(engine):
    taps = np.hanning(33) / np.hanning(33).sum()
    engine.set_analysis_config(taps=taps, delay="same", length="ceil", psd_scaling="spectrum", psd_detrend="constant",
                               psd_precision="f64")
    c = engine.analysis_config()
    assert np.array_equal(c["taps"], taps) and c["delay"] == "same" and c["length"] == "ceil"
    assert c["psd_scaling"] == "spectrum" and c["psd_detrend"] == "constant" and c["psd_precision"] == "f64"
    engine.reset_analysis_config()
    c = engine.analysis_config()
    assert c["taps"] is None and c["delay"] == "causal" and c["length"] == "floor" and not c["strict_reference"]
    with pytest.raises(EngineError):
        engine.set_analysis_config(taps=np.array([1.0, np.nan]))
    assert engine.analysis_config()["taps"] is None            # a rejected profile leaves the old one in place


def test_strict_reference_decodes(engine):
    """strict_reference reproduces the reference's fall-through decodes (VERDICT r01 missing 6): in the spectrogram cf64
    and unknown datatypes have no branch and decode to zeros (SpectralService.java:60-63: every bin 20 log10(1e-10) =
    -200 dB); in the downconverter cf64 is read at an 8-byte stride (ExtractDownConvertService.java:60-67,79-81) and
    unknown datatypes as cf32 (:93-96).  Without the flag both are errors / the correct cf64 decode."""
    from spectral_analyzer_b200 import SpectralService
    raw64 = synth.recording(4096, "cf64_le", seed=3)
    rows = engine.spectrogram(raw64, "cf64_le", 1024, 5, strict_reference=True, precision="f32")
    assert (rows[:4] == -200.0).all() and (rows[4] == -150.0).all()     # 4096 samples x 16 B: 4 readable frames
    raw32 = synth.recording(8192, "cf32_le", seed=3)
    rows = engine.spectrogram(raw32, "ri16_le", 1024, 9, strict_reference=True, out_kind="f64")
    assert (rows[:8] == -200.0).all() and (rows[8] == -150.0).all()     # Global.getBytesPerSample falls back to 8
    with pytest.raises(EngineError):
        engine.spectrogram(raw32, "ri16_le", 1024, 1)
    engine.set_analysis_config(strict_reference=True)
    try:
        assert (SpectralService(engine).computeMagnitudes(raw32, 0, 1024, "ri16_le") == -200.0).all()
        assert np.array_equal(SpectralService(engine).computeMagnitudes(raw64, 0, 256, "cf64_le"),
                              co.compute_magnitudes(raw64, 0, 256, "cf64_le", strict_reference=True))
        cfg = co.analysis_cfg(strict_reference=True)
        for dt, raw in (("cf64_le", raw64), ("cf64_be", synth.recording(4096, "cf64_be", seed=3)), ("ri16_le", raw32)):
            ref = co.downconvert_ex(raw, dt, 10, 4000, 0.05, 4, False, cfg)
            got = engine.downconvert(raw, dt, 10, 4000, 0.05, 4, False)
            assert got.shape == ref.shape and rel_err(got, ref) < DC_TOL, dt
        with pytest.raises(EngineError) as ei:
            engine.downconvert(raw64, "cf64_le", 0, 8192, 0.0, 4)          # 8192 doubles: d[i+1] of the last pair is past the buffer
        assert ei.value.code == 3
        engine.downconvert(raw64, "cf64_le", 0, 8191, 0.0, 4)
    finally:
        engine.reset_analysis_config()
    correct = engine.downconvert(raw64, "cf64_le", 10, 4000, 0.05, 4, False)
    assert rel_err(correct, co.downconvert(raw64, "cf64_le", 10, 4000, 0.05, 4, False)) < DC_TOL


def test_psd_only_batch_matches_the_batch_with_rows(engine):
    """out_iq == NULL: the decimated IQ lives only in the engine's FP32 rows; the PSD must equal the one computed in
    the same call that also returns the rows (same kernels, same rows), also across several batches / both streams."""
    raw = synth.recording(1 << 21, "ci16_le", seed=21)
    rng = np.random.default_rng(8)
    anns = [(int(rng.integers(0, (1 << 21) - 400000)), int(rng.integers(150000, 400000)), float(rng.uniform(-0.4, 0.4)), 4, False)
            for _ in range(40)]                                      # 40 x ~70k outputs x 8 B: several 24 MB batches
    _, psd_a = engine.downconvert_psd_batch(raw, "ci16_le", 1e6, anns, psd_nfft=4096, want_iq=True)
    none, psd_b = engine.downconvert_psd_batch(raw, "ci16_le", 1e6, anns, psd_nfft=4096, want_iq=False)
    assert none is None and np.array_equal(psd_a, psd_b)
    assert engine.last_kernel in ("downconvert_rows_kernel+welch_accum_mid_kernel<float,4096>",           # sa_last_kernel_name
                                  "downconvert_kernel(pipelined)+welch_accum_mid_kernel<float,4096>")
    ref = co.psd_welch(co.downconvert(raw, "ci16_le", *anns[7][:4], False), 1e6 / 4, 4096)
    strong = ref[1] > ref[1].max() - 40
    assert np.abs(psd_b[7] - ref[1])[strong].max() < 5e-3


def test_batch_matches_single_calls(engine):
    """One batched call == the per-annotation loop of AnnotationController.java:321-360."""
    raw = synth.recording(1 << 19, "cf32_le", seed=13)
    rng = np.random.default_rng(5)
    anns = []
    for i in range(12):
        down = [16, 16, 8, 32, 5][i % 5]
        count = int(rng.integers(8192 * down // 2, 12000 * down)) if i != 4 else 777 * down + 3      # one short annotation
        count = min(count, (1 << 19) - 1000)
        start = int(rng.integers(0, (1 << 19) - count))
        anns.append((start, count, float(rng.uniform(-0.4, 0.4)), down, i % 3 == 0))
    iqs, psd = engine.downconvert_psd_batch(raw, "cf32_le", 1e6, anns, psd_nfft=2048)
    for i, (s, c, f, d, fast) in enumerate(anns):
        ref = co.downconvert(raw, "cf32_le", s, c, f, d, fast)
        assert iqs[i].shape == ref.shape
        assert rel_err(iqs[i], ref) < DC_TOL
        # the Java caller's rule (AnalysisDialogController.java:303-307): nfft = min(2048, M); shorter rows are NaN padded
        nf = min(2048, ref.shape[1])
        rp = co.psd_welch(ref, 1e6 / d, nf)
        # chained FP32 stages (downconvert 1e-5 of full scale, then the FFT): compare the top 40 dB
        strong = rp[1] > rp[1].max() - 40
        assert np.abs(psd[i][:nf] - rp[1])[strong].max() < 5e-3
        assert np.isnan(psd[i][nf:]).all()


def test_golden_analysis(engine):
    g = np.load(os.path.join(GOLD, "analysis_mini.npz"))
    dc = engine.downconvert(g["raw"], "cf32_le", 100, 36000, 0.125, 4, False)
    dcf = engine.downconvert(g["raw"], "cf32_le", 100, 36000, 0.125, 4, True)
    assert rel_err(dc, g["dc"]) < DC_TOL and rel_err(dcf, g["dcf"]) < DC_TOL
    psd = engine.psd_welch(g["dc"], 1e6 / 4, 2048)
    strong = g["psd"][1] > g["psd"][1].max() - 60
    assert np.abs(psd[1] - g["psd"][1])[strong].max() < PSD_TOL_DB


def test_analysis_edge_cases(engine, tmp_path):
    """Empty and degenerate requests of the new entry points: no annotations, annotations without output, valid mode on a
    span shorter than the filter, PSD-only batches with rows that cannot hold a transform, empty file requests."""
    raw = synth.recording(50_000, "ci16_le", seed=33)
    iq, psd = engine.downconvert_psd_batch(raw, "ci16_le", 1e6, [], psd_nfft=256)
    assert iq == [] and psd.shape == (0, 256)
    # count < down: no output sample; PSD row all NaN; neighbours unaffected
    anns = [(100, 7, 0.1, 16, False), (200, 40_000, -0.2, 16, False), (0, 0, 0.0, 4, True), (5, 1000, 0.3, 4, False)]
    iq, psd = engine.downconvert_psd_batch(raw, "ci16_le", 1e6, anns, psd_nfft=256)
    assert iq[0].shape == (2, 0) and iq[2].shape == (2, 0) and np.isnan(psd[0]).all() and np.isnan(psd[2]).all()
    ref1 = co.downconvert(raw, "ci16_le", 200, 40_000, -0.2, 16, False)
    assert rel_err(iq[1], ref1) < DC_TOL and np.isfinite(psd[1]).all()
    ref3 = co.downconvert(raw, "ci16_le", 5, 1000, 0.3, 4, False)
    assert iq[3].shape == (2, 250) and rel_err(iq[3], ref3) < DC_TOL
    rp = co.psd_welch(ref3, 1e6 / 4, 250)                        # 250 < 256: one window of 250 points, rest NaN
    top = rp[1] > rp[1].max() - 40
    assert np.abs(psd[3][:250] - rp[1])[top].max() < 5e-3 and np.isnan(psd[3][250:]).all()
    _, psd_only = engine.downconvert_psd_batch(raw, "ci16_le", 1e6, anns, psd_nfft=256, want_iq=False)
    assert np.array_equal(np.isnan(psd), np.isnan(psd_only)) and np.array_equal(psd[1], psd_only[1])
    # valid mode on a span shorter than the filter: zero outputs, not an error
    engine.set_analysis_config(delay="valid")
    try:
        assert engine.downconvert_length(100, 16) == 0
        assert engine.downconvert(raw, "ci16_le", 0, 100, 0.1, 16).shape == (2, 0)
        assert engine.downconvert(raw, "ci16_le", 0, 129, 0.1, 16).shape == (2, 1)       # exactly L = 129 samples: one output
    finally:
        engine.reset_analysis_config()
    # file entry point: zero frames is a no-op, an offset past the end is an error, EOF-only requests are fill rows
    path = tmp_path / "r.sigmf-data"
    raw.tofile(path)
    assert engine.spectrogram_file(str(path), "ci16_le", 1024, 0).shape == (0, 1024)
    with pytest.raises(EngineError) as ei:
        engine.spectrogram_file(str(path), "ci16_le", 1024, 1, data_offset=10**9)
    assert ei.value.code == 3
    rows = engine.spectrogram_file(str(path), "ci16_le", 1024, 3, start_sample=49_500)
    assert (rows == -150.0).all()


def test_batch_of_more_annotations_than_one_launch_holds(engine):
    """70 000 tiny annotations: more than the 65 535 a launch can index (grid.y); the batch is cut, results are the same
    as for the first and the last annotation alone."""
    raw = synth.recording(80_000, "cu8", seed=5)
    anns = [(i, 64, 0.01 * (i % 7), 4, False) for i in range(70_000)]
    iq, _ = engine.downconvert_psd_batch(raw, "cu8", 1e6, anns, want_psd=False)
    assert len(iq) == 70_000 and all(z.shape == (2, 16) for z in (iq[0], iq[-1]))
    for i in (0, 65_534, 65_535, 69_999):
        ref = co.downconvert(raw, "cu8", *anns[i][:4], False)
        assert np.abs(iq[i] - ref).max() <= DC_TOL * max(np.abs(ref).max(), 0.5)


@pytest.mark.parametrize("dt", ["cf32_le", "cf32_be", "ci16_le", "ci16_be", "cu8", "ci8"])
@pytest.mark.parametrize("down", [8, 16, 32, 6, 10, 12, 24, 30])
@pytest.mark.parametrize("start,delay", [(0, "causal"), (1, "causal"), (122, "same"), (123, "same"), (4096, "valid"), (4097, "valid")])
def test_row_kernel_alignment_edges_and_tiles(engine, dt, down, start, delay):
    """Power-of-two decimation takes the row-per-thread kernel (16-byte cp.async of raw rows): both parities of
    (start + delay shift) -- the two decompositions of the tap sum --, an annotation that begins at sample 0 of the
    recording and one that ends at its last sample (chunks outside the recording are not read), zero history / zero
    tail inside a longer recording, several tiles per CTA, a count that is not a multiple of down."""
    count = 2 * 248 * down * 5 + 3 * down + 5            # a little over 10 tiles
    for tail in (0, 1, 333):                             # annotation ends at the end of the buffer / before it
        raw = synth.recording(start + count + tail, dt, seed=21 + tail)
        cfg = co.analysis_cfg(delay=delay, length="ceil")
        ref = co.downconvert_ex(raw, dt, start, count, -0.3217, down, False, cfg)
        engine.set_analysis_config(delay=delay, length="ceil")
        try:
            got = engine.downconvert(raw, dt, start, count, -0.3217, down, False)
            name = engine.last_kernel
        finally:
            engine.reset_analysis_config()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= DC_TOL * max(np.abs(ref).max(), 0.5), (tail, name)
        spc = 2 if dt.startswith("cf32") else 4 if dt.startswith("ci16") else 8       # samples per 16-byte chunk
        if os.environ.get("SA_DC_ROWS", "1") != "0":
            # rows of whole chunks take the row kernel (the host call packs spans on chunk boundaries and the delay shifts
            # of the built-in taps, 0 / 4 down / 8 down, are multiples of every chunk size here); the rest the staged one
            assert ("downconvert_rows_kernel" in name) == (down % spc == 0), (name, down, spc)


@pytest.mark.parametrize("count", [1, 15, 16, 17, 100, 248 * 16, 248 * 16 + 1, 5000])
def test_row_kernel_short_annotations(engine, count):
    """Annotations shorter than one tile, shorter than the filter, down to one sample; caller taps shorter than 8D+1."""
    raw = synth.recording(9000, "cf32_le", seed=33)
    taps = np.hamming(65) / np.hamming(65).sum()
    for tp in (None, taps):
        for start in (0, 2001, 9000 - count):
            cfg = co.analysis_cfg(taps=tp, delay="same", length="ceil")
            ref = co.downconvert_ex(raw, "cf32_le", start, count, 0.0613, 16, False, cfg)
            engine.set_analysis_config(taps=tp, delay="same", length="ceil")
            try:
                got = engine.downconvert(raw, "cf32_le", start, count, 0.0613, 16, False)
            finally:
                engine.reset_analysis_config()
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= DC_TOL * max(np.abs(ref).max(), 0.5), (start, tp is None)


@pytest.mark.parametrize("dt,code", [("cf32_le", 0), ("ci16_le", 1), ("cu8", 2)])
def test_row_kernel_device_path_odd_starts_and_recording_ends(engine, dt, code):
    """The device-resident batch call keeps the caller's sample offsets (the host call packs every span at an even
    offset): odd and even starts (for the integer types: every residue of the start modulo the samples per 16-byte
    chunk -- two of them take the row kernel, the others the staged kernel), annotations touching sample 0 and the last
    sample of the recording, mixed decimation factors in one call, each against the oracle."""
    import ctypes as C
    import torch
    from spectral_analyzer_b200 import _capi
    n = 300001
    raw = synth.recording(n, dt, seed=17)
    d_raw = torch.from_numpy(np.frombuffer(raw, np.uint8).copy()).cuda()
    specs = [(0, 100000, 16), (1, 100001, 16), (12345, 77777, 16), (200000, 100001, 16), (199999, 100002, 8),
             (3, 299998, 32), (150001, 40000, 16), (150002, 40000, 16), (150003, 40000, 16), (150004, 40000, 16),
             (150005, 40000, 8), (150006, 40000, 8), (150007, 40000, 32), (150008, 40000, 32),
             (5, 60000, 10), (6, 60001, 12), (7, 60002, 24), (8, 60003, 30), (11, 60000, 6), (12, 60000, 20)]
    fast_from = len(specs)                                   # box-car mode: chunk-aligned starts take the row kernel, the others the staged one
    specs += [(16, 50000, 16), (17, 50000, 16), (24, 50001, 6), (3, 50001, 8), (299000, 1001, 4)]
    anns = (_capi.Annotation * len(specs))()
    offs = (C.c_uint64 * len(specs))()
    total = 0
    for i, (s, c, d) in enumerate(specs):
        anns[i] = _capi.Annotation(s, c, 0.05 + 0.031 * i, d, 1 if i >= fast_from else 0)
        offs[i] = total
        total += 2 * (c // d)
    d_iq = torch.empty(total, dtype=torch.float64, device="cuda")
    _capi.check(_capi.lib().sa_downconvert_psd_batch_device(
        engine.handle, d_raw.data_ptr(), d_raw.numel(), code, 0, 1.0e6, anns, len(specs), 0, 0, 1, d_iq.data_ptr(), offs,
        None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got_all = d_iq.cpu().numpy()
    for i, (s, c, d) in enumerate(specs):
        ref = co.downconvert(raw, dt, s, c, 0.05 + 0.031 * i, d, i >= fast_from)
        got = got_all[offs[i]:offs[i] + 2 * (c // d)].reshape(2, c // d)
        assert rel_err(got, ref) < DC_TOL, (i, s, c, d)


@pytest.mark.parametrize("ntaps,delay", [(127, "same"), (128, "valid"), (129, "same"), (31, "same")])
def test_row_kernel_odd_delay_shift(engine, ntaps, delay):
    """The host call packs spans at even offsets, so the parity of the delay shift alone picks the decomposition:
    (L-1)/2 = 63 and L-1 = 127 are odd (rows end at the output's sample), 64 and 15 exercise the other one."""
    down, count = 16, 50003
    raw = synth.recording(count + 10, "cf32_le", seed=41)
    k = np.arange(ntaps) - (ntaps - 1) / 2
    taps = np.sinc(k / down) * np.hamming(ntaps)
    taps /= taps.sum()
    cfg = co.analysis_cfg(taps=taps, delay=delay, length="floor")
    ref = co.downconvert_ex(raw, "cf32_le", 5, count, -0.0917, down, False, cfg)
    engine.set_analysis_config(taps=taps, delay=delay, length="floor")
    try:
        got = engine.downconvert(raw, "cf32_le", 5, count, -0.0917, down, False)
    finally:
        engine.reset_analysis_config()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= DC_TOL * max(np.abs(ref).max(), 0.5)


@pytest.mark.parametrize("sched", ["0", "1"])
def test_several_batches_on_two_streams(sched):
    """The batch cut (SA_DC_BATCH_MB, read once per process) is far above these sizes by default, so the multi-batch
    path -- rows in two halves of one scratch buffer, batches alternating between two streams (SA_DC_SCHED=0) or
    pipelined behind a high-priority Welch stream (=1) -- runs here in a child process with a 1 MB cut: PSD-only and
    rows + PSD results must be identical to each other and agree with the oracle."""
    import subprocess
    import sys
    code = r'''
import numpy as np
from oracle import c_oracle as co
from spectral_analyzer_b200 import synth, Engine
eng = Engine()
raw = synth.recording(1 << 21, "cf32_le", seed=21)
rng = np.random.default_rng(8)
anns = [(int(rng.integers(0, (1 << 21) - 400000)), int(rng.integers(150000, 400000)), float(rng.uniform(-0.4, 0.4)), [4, 16, 5][i % 3], False)
        for i in range(24)]
iqs, psd_a = eng.downconvert_psd_batch(raw, "cf32_le", 1e6, anns, psd_nfft=4096, want_iq=True)
none, psd_b = eng.downconvert_psd_batch(raw, "cf32_le", 1e6, anns, psd_nfft=4096, want_iq=False)
assert none is None and np.array_equal(psd_a, psd_b)
for i in (0, 7, 13, 23):
    s, c, f, d, fast = anns[i]
    ref = co.downconvert(raw, "cf32_le", s, c, f, d, fast)
    assert np.abs(iqs[i] - ref).max() / np.abs(ref).max() < 1e-5
    rp = co.psd_welch(ref, 1e6 / d, 4096)
    strong = rp[1] > rp[1].max() - 40
    assert np.abs(psd_b[i] - rp[1])[strong].max() < 5e-3
print("ok")
'''
    env = dict(os.environ, SA_DC_BATCH_MB="1", SA_DC_SCHED=sched)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


@pytest.mark.parametrize("dt", ["cf32_le", "cf32_be", "ci16_le", "cu8", "ci8"])
@pytest.mark.parametrize("down", [2, 4, 6, 8, 16, 30, 40, 44, 64, 7])
def test_row_kernel_boxcar_mode(engine, dt, down):
    """fast = moving average then decimate: the mean of one row.  Decimation factors that fill whole 16-byte chunks (and a
    tile that fits shared memory) take downconvert_rows_fast_kernel, the others the staged kernel; several tiles, a count
    that is not a multiple of down with the ceil rule (the last row is cut by the annotation end), the annotation at the
    start and at the end of the recording."""
    count = 128 * down * 5 + 3 * down + 1
    for start, tail in ((0, 0), (64, 5), (24, 0)):
        raw = synth.recording(start + count + tail, dt, seed=50 + tail)
        for length in ("floor", "ceil"):
            cfg = co.analysis_cfg(length=length)
            ref = co.downconvert_ex(raw, dt, start, count, 0.2317, down, True, cfg)
            engine.set_analysis_config(length=length)
            try:
                got = engine.downconvert(raw, dt, start, count, 0.2317, down, True)
                name = engine.last_kernel
            finally:
                engine.reset_analysis_config()
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= DC_TOL * max(np.abs(ref).max(), 0.5), (start, length, name)
            spc = 2 if dt.startswith("cf32") else 4 if dt.startswith("ci16") else 8
            fits = 128 + 2 * 128 * ((down // spc) | 1) * 16 <= 72 * 1024
            if os.environ.get("SA_DC_ROWS", "1") != "0":
                assert (name == "downconvert_rows_fast_kernel") == (down % spc == 0 and fits), (name, down)
