#!/usr/bin/env python
"""BASELINE.json config 4 as a multi-GPU run: cu8 RTL-SDR-style recording, 2^30 samples per GPU (16 GiB over 8
GPUs), 2048-pt spectrogram fused with colormap-to-RGBA, time-sharded (contiguous block per rank, no collective on
the data path), plus the display assembly the app needs: every rank renders its columns of a canvas from its own
block (max pooling) and the canvas-sized tiles are gathered on rank 0 over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_c4_sharded.py

Supplementary to bench.py (which owns the headline line); prints one JSON object on rank 0.
"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import spectral_analyzer_b200 as sa                      # noqa: E402
from spectral_analyzer_b200 import _capi, sharding       # noqa: E402
from bench import hbm_peak                               # noqa: E402

NFFT, LOG2_PER_GPU, STEPS = 2048, 30, 10
CANVAS_W, CANVAS_H = 2048, 1024


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << LOG2_PER_GPU
    frames = n // NFFT                                   # hop = nfft: no halo (reference framing)
    g = torch.Generator(device=dev)
    g.manual_seed(4 + rank)
    raw = torch.randint(0, 256, (2 * n,), device=dev, dtype=torch.uint8, generator=g)
    out = torch.empty(frames * NFFT, device=dev, dtype=torch.int32)
    eng = sa.Engine(local)
    p = eng.make_params("cu8", NFFT, NFFT, "rect", n_frames=frames, out="rgba8", colormap="Heatmap", sample_rate=2.4e6)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.spectrogram_device(raw.data_ptr(), raw.numel(), p, out.data_ptr(), out.numel() * 4, stream)
    for _ in range(3):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / STEPS], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # display assembly: this rank's columns of the whole-recording canvas, then one gather of canvas-sized tiles
    c0, c1, _, _ = sharding.canvas_columns(CANVAS_W, 1, world, rank)
    cols = c1 - c0
    fpc = frames // max(cols, 1)
    tile = torch.empty((CANVAS_H, cols), device=dev, dtype=torch.int32)
    pc = eng.make_params("cu8", NFFT, NFFT, "rect", colormap="Heatmap", sample_rate=2.4e6)
    L = _capi.lib()

    def render():
        _capi.check(L.sa_render_canvas_device(eng.handle, raw.data_ptr(), raw.numel(), C.byref(pc), cols, CANVAS_H, fpc,
                                              _capi.REDUCE["max"], tile.data_ptr(), stream))
    tile4 = tile.view(torch.uint8).view(CANVAS_H, cols, 4)
    render()                                             # warm-up incl. the lazy NCCL communicator setup
    if world > 1:
        sharding.gather_canvas(tile4, CANVAS_W, dst=0)
    barrier()
    e0.record()
    render()
    canvas = sharding.gather_canvas(tile4, CANVAS_W, dst=0) if world > 1 else tile4
    e1.record()
    barrier()
    tg = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    if rank == 0:
        alg = world * (2 * n + frames * NFFT * 4)
        peak, kind = hbm_peak()
        print(json.dumps({
            "config": "C4 cu8 2048 rect RGBA heatmap, 2^%d samples per GPU (%.0f GiB recording), time-sharded" % (LOG2_PER_GPU, world * 2 * n / 2 ** 30),
            "n_gpus": world, "ms_per_step": round(ms, 4), "Msamples_per_s": round(world * n / ms / 1e3, 1),
            "GBps_aggregate": round(alg / ms / 1e6, 1), "roofline_frac_per_gpu": round(alg / world / ms / 1e6 / peak, 4),
            "canvas": "%dx%d max-pooled, rendered per rank + NCCL gather of tiles" % (CANVAS_W, CANVAS_H),
            "canvas_render_plus_gather_ms": round(float(tg.item()), 3), "canvas_bytes_gathered": CANVAS_W * CANVAS_H * 4,
            "canvas_shape": list(canvas.shape), "peak_kind": kind}))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
