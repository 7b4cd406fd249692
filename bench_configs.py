#!/usr/bin/env python
"""bench_configs.py -- device-resident throughput of the other BASELINE.json configurations (not the
headline line; bench.py owns that).  Prints one JSON object per configuration with the achieved fraction
of the HBM roofline from the algorithmic bytes of SURVEY.md 8d.  Synthetic data, one GPU.

    python bench_configs.py [--scale 1.0]     # --scale shrinks the sample counts (smoke runs)
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import spectral_analyzer_b200 as sa                      # noqa: E402
from spectral_analyzer_b200 import _capi                 # noqa: E402
from bench import hbm_peak                               # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    return t[len(t) // 2]


def spectrogram_case(eng, name, datatype, n_samples, nfft, hop, window, out_kind, steps, byte_offset=0, **kw):
    """byte_offset: the capture starts that many bytes into the device buffer (a WAV file's 44-byte header:
    frames are then only element-aligned and the engine takes its any-alignment kernels)."""
    kind = datatype.split("_")[0]
    bps = {"cf32": 8, "ci16": 4, "cu8": 2, "ci8": 2, "cf64": 16}[kind]
    dev = torch.device("cuda", torch.cuda.current_device())
    pad = (byte_offset + bps - 1) // bps
    if kind == "cf32":
        raw = torch.randn(2 * (n_samples + pad), device=dev, dtype=torch.float32).mul_(0.1)
    elif kind == "cf64":
        raw = torch.randn(2 * (n_samples + pad), device=dev, dtype=torch.float64).mul_(0.1)
    elif kind == "ci16":
        raw = torch.randint(-20000, 20000, (2 * (n_samples + pad),), device=dev, dtype=torch.int16)
    else:
        raw = torch.randint(0, 255, (2 * (n_samples + pad),), device=dev, dtype=torch.uint8)
    frames = (n_samples - nfft) // hop + 1
    obytes = {"f32": 4, "f64": 8, "rgba8": 4}[out_kind]
    out = torch.empty(frames * nfft * obytes, device=dev, dtype=torch.uint8)
    p = eng.make_params(datatype, nfft, hop, window, n_frames=frames, out=out_kind, **kw)
    stream = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: eng.spectrogram_device(raw.data_ptr() + byte_offset, n_samples * bps, p, out.data_ptr(), out.numel(), stream), steps)
    alg = n_samples * bps + frames * nfft * obytes
    peak, kind_p = hbm_peak()
    # the FP32 (FP64) pipe beside the HBM roofline: 5 nfft log2(nfft) / hop + 6 flop per input sample
    # against 148 SMs x 128 (64) FMA lanes x 2 flop at 1965 MHz -- for the integer inputs and the large transforms
    # the arithmetic, not the bytes, sets the time
    import math
    flops = frames * hop * (5.0 * nfft * math.log2(nfft) / hop + 6.0)
    pipe_peak = 148 * (64 if kind == "cf64" or kw.get("precision") == "f64" else 128) * 2 * 1.965e9
    res = {"config": name, "datatype": datatype, "samples": n_samples, "nfft": nfft, "hop": hop, "window": window,
           "out": out_kind, "ms": round(ms, 4), "Msamples_per_s": round(frames * hop / ms / 1e3, 1),
           "alg_bytes": alg, "GBps": round(alg / ms / 1e6, 1), "roofline_frac": round(alg / ms / 1e6 / peak, 4),
           "fft_flop_pipe_frac": round(flops / (ms * 1e-3) / pipe_peak, 4),
           "peak_kind": kind_p, "kernel": eng.last_kernel}
    del raw, out
    torch.cuda.empty_cache()
    return res


def annotation_case(eng, n_samples, n_ann, count, down, steps, want_iq=True):
    """want_iq=False: only the PSD rows are asked for (the Analysis dialog's PSD tab): the decimated IQ never
    leaves the chip and the algorithmic bytes drop by its 16 B per output sample."""
    dev = torch.device("cuda", torch.cuda.current_device())
    raw = torch.randn(2 * n_samples, device=dev, dtype=torch.float32).mul_(0.1)
    rng = np.random.default_rng(3)
    anns = (_capi.Annotation * n_ann)()
    offs = (C.c_uint64 * n_ann)()
    m = count // down
    for i in range(n_ann):
        anns[i] = _capi.Annotation(int(rng.integers(0, n_samples - count)), count, float(rng.uniform(-0.4, 0.4)), down, 0)
        offs[i] = i * 2 * m
    out_iq = torch.empty(n_ann * 2 * m, device=dev, dtype=torch.float64)
    out_psd = torch.empty(n_ann * 8192, device=dev, dtype=torch.float64)
    stream = torch.cuda.current_stream().cuda_stream
    L = _capi.lib()

    def run():
        _capi.check(L.sa_downconvert_psd_batch_device(eng.handle, raw.data_ptr(), n_samples * 8, 0, 0, 1.0e6, anns, n_ann,
                                                      8192, 2048, 1, out_iq.data_ptr() if want_iq else None,
                                                      offs if want_iq else None, out_psd.data_ptr(), stream))
    ms = timed(run, steps)
    alg = n_ann * (8 * count + (16 * m if want_iq else 0) + 8 * 8192)
    peak, kind_p = hbm_peak()
    return {"config": "C3 %d annotations x 2^%d cf32 samples: NCO + FIR /%d + Welch 8192 (75%% overlap)%s"
                      % (n_ann, count.bit_length() - 1, down, "" if want_iq else ", PSD only (no decimated IQ written)"),
            "annotations": n_ann, "count": count, "down": down, "psd_nfft": 8192,
            "ms": round(ms, 4), "Msamples_per_s": round(n_ann * count / ms / 1e3, 1), "alg_bytes": alg,
            "GBps": round(alg / ms / 1e6, 1), "roofline_frac": round(alg / ms / 1e6 / peak, 4), "peak_kind": kind_p,
            "kernel": eng.last_kernel}


def canvas_case(eng, n_samples, nfft, hop, W, H, reduce, steps, host=True):
    """N2: whole-recording view, canvas W x H from 2^28 cf32 samples: device-resident time of the spectrogram +
    canvas kernels, and the host call (H2D of the samples, D2H of the canvas only)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    raw = torch.randn(2 * n_samples, device=dev, dtype=torch.float32).mul_(0.1)
    frames = (n_samples - nfft) // hop + 1
    fpc = frames // W
    out = torch.empty(W * H, device=dev, dtype=torch.int32)
    p = eng.make_params("cf32_le", nfft, hop, "hann", colormap="Heatmap", sample_rate=2.4e6)
    stream = torch.cuda.current_stream().cuda_stream
    L = _capi.lib()
    red = _capi.REDUCE[reduce]

    def run():
        _capi.check(L.sa_render_canvas_device(eng.handle, raw.data_ptr(), n_samples * 8, C.byref(p), W, H, fpc, red,
                                              out.data_ptr(), stream))
    ms = timed(run, steps)
    alg = (W * nfft if reduce == "nearest" else n_samples) * 8 + W * H * 4
    peak, kind_p = hbm_peak()
    res = {"config": "N2 canvas %dx%d (%s) from cf32 %d-pt hop %d" % (W, H, reduce, nfft, hop), "samples": n_samples,
           "frames_per_column": fpc, "ms": round(ms, 4), "Msamples_per_s": round(W * fpc * hop / ms / 1e3, 1),
           "alg_bytes": alg, "GBps": round(alg / ms / 1e6, 1), "roofline_frac": round(alg / ms / 1e6 / peak, 4),
           "d2h_bytes": W * H * 4, "peak_kind": kind_p, "kernel": eng.last_kernel}
    if not host:
        return res
    h_raw = torch.empty(2 * n_samples, dtype=torch.float32, pin_memory=True)
    h_raw.copy_(raw)
    h_np = h_raw.numpy()
    import time
    eng.render_canvas(h_np, "cf32_le", nfft, W, H, 2.4e6, hop=hop, window="hann", frames_per_column=fpc, reduce=reduce,
                      colormap="Heatmap")
    t0 = time.perf_counter()
    for _ in range(3):
        eng.render_canvas(h_np, "cf32_le", nfft, W, H, 2.4e6, hop=hop, window="hann", frames_per_column=fpc,
                          reduce=reduce, colormap="Heatmap")
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    res.update({"host_call_ms": round(host_ms, 2), "host_Msamples_per_s": round(W * fpc * hop / host_ms / 1e3, 1)})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="", help="substring filter on the configuration name")
    args = ap.parse_args()
    sc = lambda n: max(1 << 20, int(n * args.scale))
    eng = sa.Engine(0)
    res = []
    cases = []
    cases.append(('C1-parity (reference framing)', lambda: spectrogram_case(eng, "C1-parity (reference framing)", "cf32_le", sc(1 << 28), 1024, 1024, "rect", "f32", args.steps)))
    cases.append(('C2 ci16 4096 Blackman-Harris', lambda: spectrogram_case(eng, "C2 ci16 4096 Blackman-Harris", "ci16_le", sc(1 << 30), 4096, 4096, "blackman_harris", "f32", args.steps)))
    cases.append(('C2b ci16 4096 Blackman-Harris 50% overlap', lambda: spectrogram_case(eng, "C2b ci16 4096 Blackman-Harris 50% overlap", "ci16_le", sc(1 << 29), 4096, 2048, "blackman_harris", "f32", args.steps)))
    cases.append(('C3 annotation analysis', lambda: annotation_case(eng, sc(1 << 29), 500, 1 << 20, 16, args.steps)))
    cases.append(("C4 cu8 2048 RGBA heatmap (1-GPU slice)", lambda: spectrogram_case(
        eng, "C4 cu8 2048 RGBA heatmap (1-GPU slice)", "cu8", sc(1 << 31), 2048, 2048, "rect", "rgba8", args.steps,
        colormap="Heatmap", sample_rate=2.4e6)))
    cases.append(("C4b cu8 2048 Hann RGBA", lambda: spectrogram_case(
        eng, "C4b cu8 2048 Hann RGBA", "cu8", sc(1 << 30), 2048, 2048, "hann", "rgba8", args.steps,
        colormap="Heatmap", sample_rate=2.4e6)))
    cases.append(("C4c cu8 2048 rect f32 dB", lambda: spectrogram_case(
        eng, "C4c cu8 2048 rect f32 dB", "cu8", sc(1 << 30), 2048, 2048, "rect", "f32", args.steps)))
    cases.append(('C5 cf64 65536 Hann f64', lambda: spectrogram_case(eng, "C5 cf64 65536 Hann f64", "cf64_le", sc(1 << 26), 65536, 65536, "hann", "f64", args.steps)))
    cases.append(('cf32 65536 Hann (four-step FP32)', lambda: spectrogram_case(eng, "cf32 65536 Hann (four-step FP32)", "cf32_le", sc(1 << 27), 65536, 65536, "hann", "f32", args.steps)))
    cases.append(('cf32 256 rect', lambda: spectrogram_case(eng, "cf32 256 rect", "cf32_le", sc(1 << 28), 256, 256, "rect", "f32", args.steps)))
    cases.append(('cf32 64 rect', lambda: spectrogram_case(eng, "cf32 64 rect", "cf32_le", sc(1 << 28), 64, 64, "rect", "f32", args.steps)))
    cases.append(('cf32 128 Hann', lambda: spectrogram_case(eng, "cf32 128 Hann", "cf32_le", sc(1 << 28), 128, 128, "hann", "f32", args.steps)))
    cases.append(('cf32 512 Hann 50% overlap', lambda: spectrogram_case(eng, "cf32 512 Hann 50% overlap", "cf32_le", sc(1 << 28), 512, 256, "hann", "f32", args.steps)))
    cases.append(('cf32 16384 Hann', lambda: spectrogram_case(eng, "cf32 16384 Hann", "cf32_le", sc(1 << 28), 16384, 16384, "hann", "f32", args.steps)))
    cases.append(("cf64 1024 Hann FP64", lambda: spectrogram_case(eng, "cf64 1024 Hann FP64", "cf64_le", sc(1 << 27), 1024, 1024, "hann", "f64", args.steps)))
    cases.append(("cf64 256 Hann FP64", lambda: spectrogram_case(eng, "cf64 256 Hann FP64", "cf64_le", sc(1 << 27), 256, 256, "hann", "f64", args.steps)))
    cases.append(("cf64 8192 Hann FP64", lambda: spectrogram_case(eng, "cf64 8192 Hann FP64", "cf64_le", sc(1 << 27), 8192, 8192, "hann", "f64", args.steps)))
    cases.append(("ci16 1024 Hann 50% overlap", lambda: spectrogram_case(eng, "ci16 1024 Hann 50% overlap", "ci16_le", sc(1 << 29), 1024, 512, "hann", "f32", args.steps)))
    cases.append(("cf32 8192 Hann", lambda: spectrogram_case(eng, "cf32 8192 Hann", "cf32_le", sc(1 << 28), 8192, 8192, "hann", "f32", args.steps)))
    cases.append(("N2 canvas max", lambda: canvas_case(eng, sc(1 << 28), 1024, 512, 2048, 1024, "max", args.steps)))
    cases.append(("N2 canvas nearest", lambda: canvas_case(eng, sc(1 << 28), 1024, 512, 2048, 1024, "nearest", args.steps)))
    for name, fn in cases:
        if args.only in name:
            res.append(fn())
    for r in res:
        print(json.dumps(r))
    eng.close()


if __name__ == "__main__":
    main()
