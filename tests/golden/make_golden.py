"""Generates tests/golden/*.npz.

The reference (Java 21 + commons-math3 + JDSP) cannot run in the build container (no JVM) and
holds no golden vectors of its own, so these fixtures come from oracle/np_oracle.py -- the numpy
FP64 restatement of the reference's arithmetic -- NOT from the reference itself (parity unpinned).
They pin the oracle pair (C and numpy) and the CUDA path against accidental drift.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import np_oracle as no                      # noqa: E402
from spectral_analyzer_b200 import synth                # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # 1. reference-parity spectrogram (rect, hop = nfft, 20log10(|X|+1e-10)), every decode branch
    for dt in ("cf32_le", "cf32_be", "ci16_le", "ci16_be", "cu8", "ci8", "cf64_le"):
        nfft, frames = 256, 6
        raw = synth.recording(nfft * frames - 40, dt, seed=11)       # last frame runs past EOF
        img = no.spectrogram(raw, dt, 0, nfft, nfft, "rect", frames)
        np.savez_compressed(os.path.join(HERE, "spec_parity_%s.npz" % dt), raw=raw, nfft=nfft, frames=frames,
                            datatype=dt, img=img.astype(np.float64))
    # 2. headline configuration in miniature: cf32, 1024-pt Hann, 50 % overlap
    raw = synth.recording(1024 * 5, "cf32_le", seed=1)
    img = no.spectrogram(raw, "cf32_le", 0, 1024, 512, "hann", 9)
    np.savez_compressed(os.path.join(HERE, "spec_c1_mini.npz"), raw=raw, nfft=1024, hop=512, frames=9, img=img)
    # 3. colormap
    rgba_h = no.render_rgba(img, 2.4e6, -160.0, -30.0, "Heatmap")
    rgba_g = no.render_rgba(img, 2.4e6, -160.0, -30.0, "Grayscale")
    np.savez_compressed(os.path.join(HERE, "render_c1_mini.npz"), heatmap=rgba_h, grayscale=rgba_g, fs=2.4e6)
    # 4. downconvert + Welch (self-defined spec)
    raw = synth.recording(40000, "cf32_le", seed=3)
    dc = no.downconvert(raw, "cf32_le", 100, 36000, 0.125, 4, False)
    dcf = no.downconvert(raw, "cf32_le", 100, 36000, 0.125, 4, True)
    psd = no.psd_welch(dc, 1e6 / 4, 2048)
    np.savez_compressed(os.path.join(HERE, "analysis_mini.npz"), raw=raw, dc=dc, dcf=dcf, psd=psd)
    # 5. IqData packers and analysis series on the decimated signal above, plus edge values
    edge = np.array([[0.0, 1.0, -1.0, 0.999985, 1.00002, -1.00004, 3.2, -7.9, 1e12, -1e12, np.nan, 1e-9],
                     [0.5, -0.5, 2.0 ** -15, -(2.0 ** -15), 0.99999, -0.99999, 65536.0, -65536.5, np.inf, -np.inf, 0.25, -0.0]])
    mag, frq = no.analysis_series(dc, 250e3, 0.2, 0.05, 915e6)
    np.savez_compressed(os.path.join(HERE, "iqdata_mini.npz"), edge=edge,
                        edge_f32=np.frombuffer(no.iq_pack(edge, "float32"), np.uint8),
                        edge_i16=np.frombuffer(no.iq_pack(edge, "int16"), np.uint8),
                        dc_i16=np.frombuffer(no.iq_pack(dc, "int16"), np.uint8), mag=mag, frq=frq)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
