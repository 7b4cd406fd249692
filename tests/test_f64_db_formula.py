"""The arithmetic of the FP64 dB epilogue (csrc/spectrogram_kernel.cuh: log2_tab / bins_to_db<double>), restated in
numpy and checked on the CPU against 20 log10(sqrt(p) + 1e-10) in 80-bit long double:

    dB = 10 log10(2) * (e + T[k] + log1p(r) / ln 2) + (20 / ln 10) (eps - eps^2 / 2),   p = 2^e * m,  m in [1, 2)
    k = top 7 mantissa bits of m,  c_k ~ 1 / (1 + (k + 0.5) / 128)  (a float),  T[k] = -log2 c_k,  r = m c_k - 1,
    log1p(r) ~ r (1 + r (-1/2 + r (1/3 - r / 4))),  eps = 1e-10 / sqrt(p) in float32,  valid for p >= 1e-10.

The device takes c_k from MUFU.RCP (rcp.approx) both when it builds T and when it forms r, so only consistency
matters: here c_k is the correctly rounded float reciprocal moved by up to 2 ulp either way."""
import numpy as np
import pytest


def fast_db(p, ulp_shift):
    hi = (p.view(np.int64) >> 32).astype(np.int64)
    k = (hi >> 13) & 127
    m = ((p.view(np.int64) & 0x000FFFFFFFFFFFFF) | (0x3FF << 52)).view(np.float64)
    mid = (1 + (np.arange(128) + 0.5) / 128).astype(np.float32)
    c = (np.float32(1) / mid).astype(np.float32)
    for _ in range(abs(ulp_shift)):
        c = np.nextafter(c, np.float32(2 if ulp_shift > 0 else 0)).astype(np.float32)
    c = c.astype(np.float64)
    tab = (-np.log2(c.astype(np.longdouble))).astype(np.float64)
    r = m * c[k] - 1.0
    q = r * (-0.25) + 1.0 / 3.0
    q = r * q - 0.5
    q = r * q + 1.0
    e = ((hi >> 20) - 1023).astype(np.float64)
    l2 = (r * q) * 1.4426950408889634074 + (e + tab[k])
    with np.errstate(over="ignore"):
        eps = (np.float32(1e-10) / np.sqrt(p.astype(np.float32))).astype(np.float32)
    corr = (np.float32(8.685889638065037) * eps * (np.float32(1) - np.float32(0.5) * eps)).astype(np.float32)
    return l2 * 3.0102999566398119521 + corr.astype(np.float64), np.abs(r).max()


@pytest.mark.parametrize("ulp_shift", [-2, 0, 2])
def test_table_log_matches_literal_form(ulp_shift):
    rng = np.random.default_rng(7)
    p = 10.0 ** rng.uniform(-10, 300, 1_000_000)
    p = np.concatenate([p, 10.0 ** rng.uniform(-10, 2, 1_000_000),
                        [1e-10, 1.0, 2.0, 4.0, 1.9999999999999998, 1.0000000000000002, 9.99e299]])
    got, rmax = fast_db(p, ulp_shift)
    ref = 20 * np.log10(np.sqrt(p.astype(np.longdouble)) + np.longdouble(1e-10))
    assert rmax <= 2.0 ** -8 + 1e-6
    assert float(np.abs(got - ref).max()) < 1e-10            # the FP64 parity tolerance is 1e-9 dB


def test_threshold_is_where_the_two_term_series_suffices():
    # eps = 1e-10 / sqrt(p) <= 1e-5 at p >= 1e-10: the dropped term (20 / ln 10) eps^3 / 3 is below 3e-15 dB
    eps = 1e-10 / np.sqrt(1e-10)
    assert eps <= 1.0000001e-5 and 8.685889638065037 * eps ** 3 / 3 < 3e-15
