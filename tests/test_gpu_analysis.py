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
        engine.psd_welch(iq, 1.0, 1000)          # not a power of two
    with pytest.raises(EngineError):
        engine.psd_welch(iq, 1.0, 2048)          # longer than the signal


def test_batch_matches_single_calls(engine):
    """One batched call == the per-annotation loop of AnnotationController.java:321-360."""
    raw = synth.recording(1 << 19, "cf32_le", seed=13)
    rng = np.random.default_rng(5)
    anns = []
    for i in range(12):
        down = [16, 16, 8, 32, 5][i % 5]
        count = int(rng.integers(8192 * down // 2, 12000 * down))
        count = min(count, (1 << 19) - 1000)
        start = int(rng.integers(0, (1 << 19) - count))
        anns.append((start, count, float(rng.uniform(-0.4, 0.4)), down, i % 3 == 0))
    iqs, psd = engine.downconvert_psd_batch(raw, "cf32_le", 1e6, anns, psd_nfft=2048)
    for i, (s, c, f, d, fast) in enumerate(anns):
        ref = co.downconvert(raw, "cf32_le", s, c, f, d, fast)
        assert iqs[i].shape == ref.shape
        assert rel_err(iqs[i], ref) < DC_TOL
        if ref.shape[1] >= 2048:
            rp = co.psd_welch(ref, 1e6 / d, 2048)
            # chained FP32 stages (downconvert 1e-5 of full scale, then the FFT): compare the top 40 dB
            strong = rp[1] > rp[1].max() - 40
            assert np.abs(psd[i] - rp[1])[strong].max() < 5e-3
        else:
            assert np.isnan(psd[i]).all()


def test_golden_analysis(engine):
    g = np.load(os.path.join(GOLD, "analysis_mini.npz"))
    dc = engine.downconvert(g["raw"], "cf32_le", 100, 36000, 0.125, 4, False)
    dcf = engine.downconvert(g["raw"], "cf32_le", 100, 36000, 0.125, 4, True)
    assert rel_err(dc, g["dc"]) < DC_TOL and rel_err(dcf, g["dcf"]) < DC_TOL
    psd = engine.psd_welch(g["dc"], 1e6 / 4, 2048)
    strong = g["psd"][1] > g["psd"][1].max() - 60
    assert np.abs(psd[1] - g["psd"][1])[strong].max() < PSD_TOL_DB
