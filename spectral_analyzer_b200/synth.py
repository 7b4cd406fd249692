"""Synthetic SigMF-style recordings for the tests and bench.py:
three complex tones (two off-bin, one on-bin) plus white Gaussian noise, SNR about 40 dB,
encoded as cf32 / ci16 / cu8 / ci8 / cf64 in either byte order.  Deterministic per seed.
"""
import numpy as np

TONES = ((0.1250, 0.5), (-0.28137, 0.25), (0.40213, 0.125))   # (cycles/sample, amplitude)
NOISE_SIGMA = 0.005
NP_DTYPE = {"cf32": "f4", "ci16": "i2", "cu8": "u1", "ci8": "i1", "cf64": "f8"}
BYTES_PER_IQ = {"cf32": 8, "ci16": 4, "cu8": 2, "ci8": 2, "cf64": 16}


def complex_signal(n, seed=1, start=0):
    rng = np.random.default_rng(seed)
    t = np.arange(start, start + n, dtype=np.float64)
    x = NOISE_SIGMA * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for f, a in TONES:
        x += a * np.exp(2j * np.pi * np.mod(f * t, 1.0))
    return x


def encode(x, datatype):
    """complex128 -> raw bytes (np.uint8 array) in the SigMF datatype, e.g. 'ci16_le'."""
    kind = datatype.split("_")[0]
    order = "<" if datatype.endswith("_le") else ">"
    iq = np.empty(2 * x.size, np.float64)
    iq[0::2], iq[1::2] = x.real, x.imag
    if kind == "cf32":
        raw = iq.astype(order + "f4")
    elif kind == "cf64":
        raw = iq.astype(order + "f8")
    elif kind == "ci16":
        raw = np.clip(np.rint(iq * 0.7 * 32767.0), -32768, 32767).astype(order + "i2")
    elif kind == "cu8":
        raw = np.clip(np.rint(iq * 0.7 * 127.5 + 127.5), 0, 255).astype("u1")
    elif kind == "ci8":
        raw = np.clip(np.rint(iq * 0.7 * 127.0), -128, 127).astype("i1")
    else:
        raise ValueError(datatype)
    return raw.view(np.uint8).reshape(-1)


def recording(n, datatype, seed=1):
    return encode(complex_signal(n, seed), datatype)
