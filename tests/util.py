"""Tolerances of the parity tests (BASELINE.json north_star):
integer decode and frame indexing bit-exact; FP32 power within 1e-5 relative on signal bins and
within 1e-3 dB on every bin above the noise floor; the FP64 path much tighter."""
import numpy as np

REL_POWER_TOL = 1e-5                     # relative power error on strong bins
DB_OF_REL_POWER = 10 * np.log10(1 + REL_POWER_TOL)   # = 4.34e-5 dB
DB_TOL_ABOVE_FLOOR = 1e-3                # dB, bins above the noise floor
STRONG_BELOW_MAX_DB = 40.0               # "signal" bins: within 40 dB of the frame maximum
FLOOR_BELOW_MAX_DB = 110.0               # bins further down are cancellation residue, compared in linear power


def check_db_parity(got, ref, mode_power=False, strong_tol=DB_OF_REL_POWER, floor_tol=DB_TOL_ABOVE_FLOOR):
    """got: engine dB image, ref: oracle FP64 dB image, same shape [frames, nfft]."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    assert np.isfinite(got).all()
    fmax = ref.max(axis=1, keepdims=True)
    diff = np.abs(got - ref)
    strong = ref >= fmax - STRONG_BELOW_MAX_DB
    above = ref >= fmax - FLOOR_BELOW_MAX_DB
    assert strong.any()
    worst_strong = diff[strong].max()
    worst_above = diff[above].max()
    assert worst_strong <= strong_tol, "strong bins: %.3g dB > %.3g dB" % (worst_strong, strong_tol)
    assert worst_above <= floor_tol, "bins above floor: %.3g dB > %.3g dB" % (worst_above, floor_tol)
    # far below the floor: compare linear amplitude against the frame maximum instead of dB
    scale = 20.0 if not mode_power else 10.0
    lin_g, lin_r = 10 ** (got / scale), 10 ** (ref / scale)
    lin_max = 10 ** (fmax / scale)
    assert (np.abs(lin_g - lin_r) <= 1e-5 * lin_max).all()
    return worst_strong, worst_above
