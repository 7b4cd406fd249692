"""numpy (pocketfft, FP64) restatement of the same path, written independently of
sa_oracle.c so that the two can check each other.  TEST INFRASTRUCTURE ONLY.
PARITY UNPINNED.  Citations: S/ = src/main/java/net/kcundercover/spectral_analyzer/.
"""
import numpy as np
import scipy.signal as sps

BYTES_PER_IQ = {"cf32": 8, "ci16": 4, "cu8": 2, "ci8": 2, "cf64": 16}   # S/sigmf/Global.java:67-79


def _split(datatype):
    kind = datatype.split("_")[0]
    order = "<" if datatype.endswith("_le") else ">"                   # S/sigmf/SigMfHelper.java:87-91
    return kind, order


def decode(buf, start_byte, count, datatype):
    """S/services/SpectralService.java:42-63, S/services/ExtractDownConvertService.java:79-98."""
    kind, order = _split(datatype)
    raw = np.frombuffer(np.ascontiguousarray(buf).view(np.uint8), np.uint8)
    nb = count * BYTES_PER_IQ[kind]
    if start_byte + nb > raw.size:
        raise IndexError("read past end of buffer")
    seg = raw[start_byte:start_byte + nb]
    if kind == "ci16":
        v = seg.view(order + "i2").astype(np.float64) / 32768.0
    elif kind == "cf32":
        v = seg.view(order + "f4").astype(np.float64)
    elif kind == "cu8":
        v = (seg.astype(np.float64) - 127.5) / 128
    elif kind == "ci8":
        v = seg.view(np.int8).astype(np.float64) / 128
    elif kind == "cf64":
        v = seg.view(order + "f8").astype(np.float64)
    else:
        v = np.zeros(2 * count)
    return v[0::2] + 1j * v[1::2]


def window(name, n):
    if name == "rect":
        return np.ones(n)
    x = 2 * np.pi * np.arange(n) / n
    if name == "hann":
        return 0.5 - 0.5 * np.cos(x)
    if name == "hamming":
        return 0.54 - 0.46 * np.cos(x)
    if name == "blackman":
        return 0.42 - 0.5 * np.cos(x) + 0.08 * np.cos(2 * x)
    if name == "blackman_harris":
        return 0.35875 - 0.48829 * np.cos(x) + 0.14128 * np.cos(2 * x) - 0.01168 * np.cos(3 * x)
    raise ValueError(name)


def compute_magnitudes(buf, start_byte, nfft, datatype):
    """S/services/SpectralService.java:33-85."""
    x = decode(buf, start_byte, nfft, datatype)
    X = np.fft.fft(x)                                                  # :68 unnormalised forward
    return np.fft.fftshift(20 * np.log10(np.abs(X) + 1e-10))           # :76-82


def spectrogram(buf, datatype, start_sample, nfft, hop, win, n_frames, db_mode=0):
    """S/controllers/MainController.java:980-999 generalised to hop/window."""
    kind, _ = _split(datatype)
    bps = BYTES_PER_IQ[kind]
    cap = np.ascontiguousarray(buf).view(np.uint8).size
    w = window(win, nfft)
    out = np.full((n_frames, nfft), -150.0)                            # :996-997
    for t in range(n_frames):
        off = (start_sample + t * hop) * bps                           # :984-985
        if off + nfft * bps <= cap:                                    # :987
            X = np.fft.fft(decode(buf, off, nfft, datatype) * w)
            if db_mode == 0:
                out[t] = np.fft.fftshift(20 * np.log10(np.abs(X) + 1e-10))
            else:
                out[t] = np.fft.fftshift(10 * np.log10(np.abs(X) ** 2 + 1e-20))
    return out


def render_rgba(db, fs, min_db, max_db, cmap):
    """S/controllers/MainController.java:1273-1285 + :926-957 (float-component Color)."""
    nfft = db.shape[-1]
    conv = 10 * np.log10(fs / nfft) + 20 * np.log10(nfft)
    n = np.clip((db - conv - min_db) / (max_db - min_db), 0.0, 1.0)

    def lerp(a, b, t):
        ft = t.astype(np.float32)
        r = np.float32(a) + (np.float32(b) - np.float32(a)) * ft
        return np.where(t <= 0, np.float32(a), np.where(t >= 1, np.float32(b), r)).astype(np.float32)

    if cmap == "Heatmap":
        u1, u2 = (n - 0.2) / 0.3, (n - 0.5) / 0.5
        lo, mid = n < 0.2, n < 0.5
        r = np.where(lo, 0, np.where(mid, lerp(0, 1, u1), 1)).astype(np.float32)
        g = np.where(mid, 0, lerp(0, 1, u2)).astype(np.float32)
        b = np.where(lo, 0, np.where(mid, lerp(1, 0, u1), 0)).astype(np.float32)
    else:
        r = g = b = lerp(0, 1, n)
    ch = lambda c: np.floor(c.astype(np.float64) * 255.0 + 0.5).astype(np.uint8)
    return np.stack([ch(r), ch(g), ch(b), np.full(n.shape, 255, np.uint8)], axis=-1)


def lowpass_taps(down):
    nt, mid, fc = 8 * down + 1, 4 * down, 0.5 / down
    k = np.arange(nt)
    h = 2 * fc * np.sinc(2 * fc * (k - mid)) * np.hamming(nt)
    return h / h.sum()


def downconvert(buf, datatype, start_sample, count, freq_off, down, fast=False):
    """Self-defined spec of S/services/ExtractDownConvertService.java:54-117 (JDSP not vendored)."""
    kind, _ = _split(datatype)
    if count // down == 0:                          # nothing to emit (the C restatement returns length 0 as well)
        return np.zeros((2, 0), np.float64)
    x = decode(buf, start_sample * BYTES_PER_IQ[kind], count, datatype)
    n = np.arange(count)
    y = x * np.exp(-2j * np.pi * np.mod(freq_off * n, 1.0))
    m = count // down
    if fast:
        return_c = y[:m * down].reshape(m, down).mean(axis=1)
    else:
        z = sps.lfilter(lowpass_taps(down), [1.0], y)
        return_c = z[::down][:m]
    return np.stack([return_c.real, return_c.imag])


def psd_welch(iq, fs, nfft, hop=None, win="hann"):
    """Self-defined spec of the JDSP call at S/controllers/AnalysisDialogController.java:308-312."""
    hop = hop or max(1, nfft // 4)
    x = np.asarray(iq[0]) + 1j * np.asarray(iq[1])
    f, p = sps.welch(x, fs=fs, window=window(win, nfft), nperseg=nfft, noverlap=nfft - hop, nfft=nfft,
                     detrend=False, return_onesided=False, scaling="density")
    return np.stack([np.fft.fftshift(f), np.fft.fftshift(10 * np.log10(p + 1e-30))])


def iq_pack(iq, fmt):
    """IqData.getInterleavedBinary (IqData.java:160-187), independent numpy restatement."""
    re, im = np.asarray(iq[0], np.float64), np.asarray(iq[1], np.float64)
    if fmt.lower() == "float32":
        out = np.empty(2 * re.size, "<f4")
        out[0::2], out[1::2] = re.astype(np.float32), im.astype(np.float32)
        return out.tobytes()
    if fmt.lower() == "int16":
        def narrow(v):
            v = 32767 * v
            i = np.where(np.isnan(v), 0.0, np.clip(np.trunc(v), -2147483648.0, 2147483647.0)).astype(np.int64)
            return (i & 0xFFFF).astype(np.uint16)
        out = np.empty(2 * re.size, "<u2")
        out[0::2], out[1::2] = narrow(re), narrow(im)
        return out.tobytes()
    raise ValueError("Unsupported binary format: " + fmt)


def analysis_series(iq, fs, alpha_mag=1.0, alpha_freq=1.0, center_freq=0.0):
    """updateMagnitudeChart / updateFrequencyChart (AnalysisDialogController.java:219-290); the EMA
    is evaluated with scipy.signal.lfilter (same recurrence, different summation order)."""
    from scipy.signal import lfilter
    re, im = np.asarray(iq[0], np.float64), np.asarray(iq[1], np.float64)

    def ema(x, alpha):
        if x.size == 0:
            return x
        y, _ = lfilter([alpha], [1.0, -(1.0 - alpha)], x[1:], zi=[(1.0 - alpha) * x[0]])
        return np.concatenate([x[:1], y])
    with np.errstate(divide="ignore"):
        mag = 20 * np.log10(ema(np.hypot(re, im), alpha_mag))
    ph = np.arctan2(im, re)
    d = ph[1:] - ph[:-1]
    d = np.where(d > np.pi, d - 2 * np.pi, np.where(d < -np.pi, d + 2 * np.pi, d))
    frq = np.full(re.size, np.nan)
    frq[1:] = ema(d / (2 * np.pi) * fs, alpha_freq) + center_freq
    return mag, frq

