"""A/B matrix for the asynchronously staged mid kernels: run twice, with SA_MID_PF=all and SA_MID_PF=none."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectral_analyzer_b200 as sa                      # noqa: E402
from bench_configs import spectrogram_case               # noqa: E402

eng = sa.Engine(0)
for nfft in [int(x) for x in os.environ.get("SA_MATRIX_N", "2048,4096,8192,16384").split(",")]:
    for dt, log2n in (("cf32_le", 28), ("ci16_le", 29), ("cu8", 30)):
        for win in ("hann", "rect"):
            r = spectrogram_case(eng, "m", dt, 1 << log2n, nfft, nfft, win, "f32", 8)
            print(json.dumps({"nfft": nfft, "dt": dt, "win": win, "ms": r["ms"], "frac": r["roofline_frac"],
                              "pf": os.environ.get("SA_MID_PF", "default")}))
eng.close()
