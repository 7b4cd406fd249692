"""ctypes front-end of oracle/libsa_oracle.so (the C FP64 restatement; see sa_oracle.h).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DTYPES = {"cf32": 0, "ci16": 1, "cu8": 2, "ci8": 3, "cf64": 4}
DT_OTHER = 5          # a datatype without a decode branch (ri16_le, cf16 ...): only meaningful with strict_reference
DELAY = {"causal": 0, "same": 1, "valid": 2}
LENGTH = {"floor": 0, "ceil": 1}
SCALING = {"density": 0, "spectrum": 1}
DETREND = {None: 0, "none": 0, False: 0, "constant": 1}


class AnalysisCfg(C.Structure):
    _fields_ = [("taps", C.POINTER(C.c_double)), ("n_taps", C.c_int), ("delay_mode", C.c_int), ("length_mode", C.c_int),
                ("psd_scaling", C.c_int), ("psd_detrend", C.c_int), ("strict_reference", C.c_int)]


def analysis_cfg(taps=None, delay="causal", length="floor", scaling="density", detrend=None, strict_reference=False):
    """ora_analysis_cfg; keeps the taps array alive on the returned object."""
    c = AnalysisCfg()
    if taps is not None:
        c._taps = np.ascontiguousarray(taps, np.float64)
        c.taps = c._taps.ctypes.data_as(C.POINTER(C.c_double))
        c.n_taps = c._taps.size
    c.delay_mode, c.length_mode = DELAY[delay], LENGTH[length]
    c.psd_scaling, c.psd_detrend, c.strict_reference = SCALING[scaling], DETREND[detrend], int(strict_reference)
    return c
WINDOWS = {"rect": 0, "hann": 1, "hamming": 2, "blackman": 3, "blackman_harris": 4}
DB_MAG_1E10, DB_POWER = 0, 1
CMAPS = {"Grayscale": 0, "Heatmap": 1}
BYTES_PER_IQ = {"cf32": 8, "ci16": 4, "cu8": 2, "ci8": 2, "cf64": 16}


def parse_datatype(datatype):
    """SigMF datatype string -> (kind, big_endian).  SigMfHelper.java:87-91: LE iff "_le"."""
    kind = datatype.split("_")[0]
    if kind not in DTYPES:
        raise ValueError("unsupported datatype " + datatype)
    return kind, (0 if datatype.endswith("_le") else 1)


def dtype_code(datatype, strict_reference=False):
    """(code, big_endian); with strict_reference a datatype without a decode branch maps to DT_OTHER."""
    kind = datatype.split("_")[0]
    be = 0 if datatype.endswith("_le") else 1
    if kind in DTYPES:
        return DTYPES[kind], be
    if strict_reference:
        return DT_OTHER, be
    raise ValueError("unsupported datatype " + datatype)


def build():
    so = os.path.join(_HERE, "libsa_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("sa_oracle.c", "sa_oracle.h", "Makefile")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, dp = C.POINTER(C.c_uint8), C.POINTER(C.c_double)
        L.ora_compute_magnitudes.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.ora_decode.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, dp, dp]
        L.ora_spectrogram.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_uint64,
                                      C.c_int, C.c_uint64, C.c_int, dp, C.c_int]
        L.ora_render_rgba.argtypes = [dp, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, u8p]
        L.ora_render_canvas.argtypes = [dp, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_int, u8p]
        L.ora_lowpass_taps.argtypes = [C.c_int, dp]
        L.ora_downconvert.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_double,
                                      C.c_int, C.c_int, dp, dp, C.POINTER(C.c_uint64)]
        L.ora_psd_welch.argtypes = [dp, dp, C.c_uint64, C.c_double, C.c_int, C.c_uint64, C.c_int, dp, dp]
        cfgp = C.POINTER(AnalysisCfg)
        L.ora_downconvert_length.restype = C.c_uint64
        L.ora_downconvert_length.argtypes = [C.c_uint64, C.c_int, C.c_int, cfgp]
        L.ora_downconvert_ex.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_double,
                                         C.c_int, C.c_int, cfgp, dp, dp, C.POINTER(C.c_uint64)]
        L.ora_psd_welch_ex.argtypes = [dp, dp, C.c_uint64, C.c_double, C.c_int, C.c_uint64, C.c_int, cfgp, dp, dp]
        L.ora_iq_pack.argtypes = [dp, dp, C.c_uint64, C.c_int, u8p]
        L.ora_analysis_series.argtypes = [dp, dp, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
        L.ora_window.argtypes = [C.c_int, C.c_int, dp]
        L.ora_fft.argtypes = [dp, dp, C.c_int]
        _LIB = L
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed rc=%d" % (what, rc))


def compute_magnitudes(buf, start_byte, nfft, datatype, strict_reference=False):
    """SpectralService.computeMagnitudes (SpectralService.java:33-85)."""
    kind, be = parse_datatype(datatype)
    a, p = _u8(buf)
    out = np.empty(nfft, np.float64)
    _chk(lib().ora_compute_magnitudes(p, a.size, start_byte, nfft, DTYPES[kind], be,
                                      int(strict_reference), _dp(out)), "compute_magnitudes")
    return out


def decode(buf, start_byte, count, datatype):
    kind, be = parse_datatype(datatype)
    a, p = _u8(buf)
    re, im = np.empty(count, np.float64), np.empty(count, np.float64)
    _chk(lib().ora_decode(p, a.size, start_byte, count, DTYPES[kind], be, 0, _dp(re), _dp(im)), "decode")
    return re, im


def spectrogram(buf, datatype, start_sample, nfft, hop, window, n_frames, db_mode=DB_MAG_1E10, nthreads=0):
    """MainController.updateDisplay frame loop (MainController.java:980-999), generalised."""
    kind, be = parse_datatype(datatype)
    a, p = _u8(buf)
    out = np.empty((n_frames, nfft), np.float64)
    _chk(lib().ora_spectrogram(p, a.size, DTYPES[kind], be, start_sample, nfft, hop, WINDOWS[window],
                               n_frames, db_mode, _dp(out), nthreads), "spectrogram")
    return out


def render_rgba(db, fs, min_db, max_db, cmap):
    db = np.ascontiguousarray(db, np.float64)
    nfr, nfft = db.shape
    out = np.empty((nfr, nfft, 4), np.uint8)
    _chk(lib().ora_render_rgba(_dp(db), nfr, nfft, fs, min_db, max_db, CMAPS[cmap],
                               out.ctypes.data_as(C.POINTER(C.c_uint8))), "render_rgba")
    return out


def render_canvas(db, canvas_h, fs, min_db, max_db, cmap):
    db = np.ascontiguousarray(db, np.float64)
    nfr, nfft = db.shape
    out = np.empty((canvas_h, nfr, 4), np.uint8)
    _chk(lib().ora_render_canvas(_dp(db), nfr, nfft, canvas_h, fs, min_db, max_db, CMAPS[cmap],
                                 out.ctypes.data_as(C.POINTER(C.c_uint8))), "render_canvas")
    return out


def lowpass_taps(down):
    t = np.empty(8 * down + 1, np.float64)
    _chk(lib().ora_lowpass_taps(down, _dp(t)), "lowpass_taps")
    return t


def downconvert(buf, datatype, start_sample, count, freq_off, down, fast=False):
    """ExtractDownConvertService.extractAndDownConvert (ExtractDownConvertService.java:54-117)."""
    kind, be = parse_datatype(datatype)
    a, p = _u8(buf)
    m = count // down
    re, im = np.empty(max(m, 1), np.float64), np.empty(max(m, 1), np.float64)
    n = C.c_uint64(0)
    _chk(lib().ora_downconvert(p, a.size, DTYPES[kind], be, start_sample, count, freq_off, down, int(fast),
                               _dp(re), _dp(im), C.byref(n)), "downconvert")
    return np.stack([re[:n.value], im[:n.value]])


def downconvert_ex(buf, datatype, start_sample, count, freq_off, down, fast=False, cfg=None):
    """The same with the JDSP unknowns as parameters (ora_analysis_cfg)."""
    cfg = cfg or analysis_cfg()
    code, be = dtype_code(datatype, bool(cfg.strict_reference))
    a, p = _u8(buf)
    m = int(lib().ora_downconvert_length(count, down, int(fast), C.byref(cfg)))
    re, im = np.empty(max(m, 1), np.float64), np.empty(max(m, 1), np.float64)
    n = C.c_uint64(0)
    _chk(lib().ora_downconvert_ex(p, a.size, code, be, start_sample, count, freq_off, down, int(fast), C.byref(cfg),
                                  _dp(re), _dp(im), C.byref(n)), "downconvert_ex")
    return np.stack([re[:n.value], im[:n.value]])


def psd_welch(iq, fs, nfft, hop=None, window="hann", cfg=None):
    """PowerSpectralDensity.calculatePsdWelch call site (AnalysisDialogController.java:303-313); any nfft >= 1."""
    re = np.ascontiguousarray(iq[0], np.float64)
    im = np.ascontiguousarray(iq[1], np.float64)
    hop = hop or max(1, nfft // 4)
    f, d = np.empty(nfft, np.float64), np.empty(nfft, np.float64)
    _chk(lib().ora_psd_welch_ex(_dp(re), _dp(im), re.size, fs, nfft, hop, WINDOWS[window],
                                C.byref(cfg) if cfg is not None else None, _dp(f), _dp(d)), "psd_welch")
    return np.stack([f, d])


def iq_pack(iq, fmt):
    """IqData.getInterleavedBinary (IqData.java:160-187)."""
    re = np.ascontiguousarray(iq[0], np.float64)
    im = np.ascontiguousarray(iq[1], np.float64)
    code = {"float32": 0, "int16": 1}[fmt.lower()]
    out = np.empty(re.size * (8 if code == 0 else 4), np.uint8)
    _chk(lib().ora_iq_pack(_dp(re), _dp(im), re.size, code, out.ctypes.data_as(C.POINTER(C.c_uint8))), "iq_pack")
    return out.tobytes()


def analysis_series(iq, fs, alpha_mag=1.0, alpha_freq=1.0, center_freq=0.0):
    """updateMagnitudeChart / updateFrequencyChart (AnalysisDialogController.java:219-290)."""
    re = np.ascontiguousarray(iq[0], np.float64)
    im = np.ascontiguousarray(iq[1], np.float64)
    mag, frq = np.empty(re.size, np.float64), np.empty(re.size, np.float64)
    _chk(lib().ora_analysis_series(_dp(re), _dp(im), re.size, fs, alpha_mag, alpha_freq, center_freq, _dp(mag),
                                   _dp(frq)), "analysis_series")
    return mag, frq


def fft(x):
    x = np.asarray(x, np.complex128)
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    _chk(lib().ora_fft(_dp(re), _dp(im), x.size), "fft")
    return re + 1j * im
