"""Arithmetic identities the FP32 epilogues rely on (csrc/spectrogram_kernel.cuh), pinned on the CPU in numpy
float32: they are claims about IEEE arithmetic, not about the GPU, so they are checked here where no GPU is needed.
The GPU parity tests check the kernels themselves against the oracle."""
import numpy as np

from oracle import np_oracle as no


def test_fp32_fast_path_threshold():
    """bins_to_db<float>: for |X|^2 >= 2^-18 (|X| >= 2^-9) the float sum |X| + 1e-10 rounds back to |X|, so
    20 log10(|X| + 1e-10) == 10 log10(|X|^2) and the square root can be skipped.  Just below the threshold the
    identity starts to fail, which is why those bins take the literal form."""
    rng = np.random.default_rng(1)
    x = (2.0 ** rng.uniform(-9, 20, 2_000_000)).astype(np.float32)
    x = np.concatenate([x, np.float32([2.0 ** -9, np.nextafter(np.float32(2.0 ** -9), np.float32(1))])])
    assert np.array_equal(x + np.float32(1e-10), x)
    below = (2.0 ** rng.uniform(-12, -10, 100_000)).astype(np.float32)
    assert not np.array_equal(below + np.float32(1e-10), below)


def test_magic_number_rounding_is_round_half_up_of_c_times_255():
    """colormap_px: round(c * 255) is read from the low mantissa byte of fma(c, 255, 1.5 * 2^23) (no F2I).  The FMA
    rounds c * 255 + 12582912 to an integer with round-half-even; the oracle uses floor(c * 255 + 0.5).  They
    differ only at exact .5 products, which a float c in [0, 1] cannot produce except c * 255 = k + 0.5 exactly."""
    rng = np.random.default_rng(2)
    c = rng.uniform(0, 1, 2_000_000).astype(np.float32)
    c = np.concatenate([c, np.float32([0.0, 1.0, 0.5, 1 / 255, 254.5 / 255, 0.2, 0.3])])
    magic = np.float32(12582912.0)
    # the fused multiply-add, exactly: double holds c * 255 + magic without rounding (24 + 8 bits), then one rounding
    fused = (c.astype(np.float64) * 255.0 + 12582912.0).astype(np.float32)
    low = (fused.view(np.uint32) & 0xFF).astype(np.int64)
    ref = np.floor(c.astype(np.float64) * 255.0 + 0.5).astype(np.int64)
    ties = np.abs((c.astype(np.float64) * 255.0) % 1.0 - 0.5) < 1e-12
    assert np.array_equal(low[~ties], ref[~ties])
    assert np.abs(low[ties] - ref[ties]).max(initial=0) <= 1
    assert fused.min() >= magic and fused.max() <= magic + 255


def test_heatmap_as_saturating_ramps_matches_the_piecewise_reference():
    """getColorForMagnitude (MainController.java:926-957) restated as the three saturating FMAs of colormap_px
    (r = sat((n - 0.2) / 0.3), g = sat(2 n - 1), b = n >= 0.2 ? 1 - r : 0) against the oracle's piecewise
    lerps: channels agree to 1 LSB everywhere (the parity tolerance), exactly away from the breakpoints."""
    n = np.linspace(0, 1, 200_001).astype(np.float32)
    sat = lambda v: np.clip(v, np.float32(0), np.float32(1)).astype(np.float32)
    r = sat(n * np.float32(1 / 0.3) + np.float32(-0.2 / 0.3))
    g = sat(n * np.float32(2) + np.float32(-1))
    b = np.where(n >= np.float32(0.2), np.float32(1) - r, np.float32(0)).astype(np.float32)
    px = np.stack([np.floor(ch.astype(np.float64) * 255 + 0.5) for ch in (r, g, b)], axis=-1).astype(np.int64)
    # the oracle takes dB rows and derives the dB/Hz conversion from the row length: db = conv + min + n (max - min)
    fs, lo, hi = 1.0e6, -160.0, -30.0
    conv = 10 * np.log10(fs / n.size) + 20 * np.log10(n.size)
    db = (conv + lo + n.astype(np.float64) * (hi - lo))[None, :]
    ref = no.render_rgba(db, fs, lo, hi, "Heatmap")[0, :, :3].astype(np.int64)
    assert np.abs(px - ref).max() <= 1
    away = (np.abs(n - 0.2) > 1e-3) & (np.abs(n - 0.5) > 1e-3)
    assert (np.abs(px - ref)[away].max(axis=1) == 0).mean() > 0.99
