"""Times the batched NCO + FIR-decimate (+ Welch PSD) call over dtype x decimation (device-resident input).
usage (GPU box): python tools/dc_matrix.py [--ann 200] [--count 1048576] [--steps 10]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_analyzer_b200 as sa  # noqa: E402
from spectral_analyzer_b200 import _capi  # noqa: E402
from bench_configs import timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ann", type=int, default=200)
    ap.add_argument("--count", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--downs", default="2,4,10,16,32")
    ap.add_argument("--fast", action="store_true", help="box-car mode (moving average, then decimate)")
    ap.add_argument("--align", type=int, default=1, help="annotation starts are multiples of this many samples (8: every "
                    "input type starts on a 16-byte boundary, as the spans packed by the host call do)")
    args = ap.parse_args()
    eng = sa.Engine(0)
    dev = torch.device("cuda", 0)
    L = _capi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    n_samples = 1 << 28
    bufs = {"cf32": (0, 8, torch.randn(2 * n_samples, device=dev, dtype=torch.float32).mul_(0.1)),
            "ci16": (1, 4, torch.randint(-3000, 3000, (2 * n_samples,), device=dev, dtype=torch.int16)),
            "cu8": (2, 2, torch.randint(0, 255, (2 * n_samples,), device=dev, dtype=torch.uint8))}
    rng = np.random.default_rng(3)
    for name, (code, bpp, raw) in bufs.items():
        for down in [int(x) for x in args.downs.split(",")]:
            n_ann, count = args.ann, args.count
            m = count // down
            anns = (_capi.Annotation * n_ann)()
            offs = (C.c_uint64 * n_ann)()
            for i in range(n_ann):
                start = int(rng.integers(0, n_samples - count)) // args.align * args.align
                anns[i] = _capi.Annotation(start, count, float(rng.uniform(-0.4, 0.4)), down, 1 if args.fast else 0)
                offs[i] = i * 2 * m
            out_iq = torch.empty(n_ann * 2 * m, device=dev, dtype=torch.float64)
            out_psd = torch.empty(n_ann * 8192, device=dev, dtype=torch.float64)

            def run():
                _capi.check(L.sa_downconvert_psd_batch_device(eng.handle, raw.data_ptr(), n_samples * bpp, code, 0, 1.0e6, anns,
                                                              n_ann, 8192, 2048, 1, out_iq.data_ptr(), offs, out_psd.data_ptr(), stream))
            ms = timed(run, args.steps)
            print(json.dumps({"dtype": name, "down": down, "fast": bool(args.fast), "align": args.align, "ms": round(ms, 4), "Gsamples_per_s": round(n_ann * count / ms / 1e6, 1),
                              "kernel": eng.last_kernel}))
            del out_iq, out_psd
    eng.close()


if __name__ == "__main__":
    main()
