"""Per-call latency of the compatibility entry points (one frame per call, like the reference's loop
MainController.java:982-999) against the batched call for the same redraw.   python tools/latency_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectral_analyzer_b200 as sa                      # noqa: E402
from spectral_analyzer_b200 import synth                 # noqa: E402

eng = sa.Engine(0)
svc = sa.SpectralService(eng)
for nfft in (1024, 4096, 65536):
    canvas_w = 1000 if nfft < 65536 else 100
    raw = synth.recording(nfft * canvas_w, "cf32_le", seed=1)
    svc.computeMagnitudes(raw, 0, nfft, "cf32_le")
    t0 = time.perf_counter()
    for t in range(canvas_w):
        svc.computeMagnitudes(raw, t * nfft * 8, nfft, "cf32_le")
    per_call = (time.perf_counter() - t0) / canvas_w
    svc.computeWaterfall(raw, 0, canvas_w, nfft, "cf32_le")
    t0 = time.perf_counter()
    for _ in range(5):
        svc.computeWaterfall(raw, 0, canvas_w, nfft, "cf32_le")
    batched = (time.perf_counter() - t0) / 5
    eng.render_canvas(raw, "cf32_le", nfft, canvas_w, 600, 2.4e6)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.render_canvas(raw, "cf32_le", nfft, canvas_w, 600, 2.4e6)
    canvas = (time.perf_counter() - t0) / 5
    print("nfft %6d canvasW %4d: computeMagnitudes %.1f us/call = %.1f ms per redraw | computeWaterfall (one call) %.2f ms | "
          "render_canvas %dx600 (one call) %.2f ms" % (nfft, canvas_w, per_call * 1e6, per_call * canvas_w * 1e3, batched * 1e3,
                                                       canvas_w, canvas * 1e3))
eng.close()
