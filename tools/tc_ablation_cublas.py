"""Tensor-core ablation (north_star: "tensor cores only if a DFT-as-GEMM stage beats the FP32 path within
tolerance").  One radix-32 DFT stage over 2^28 complex points written as the real GEMM
[2^23 x 64] . [64 x 64] (interleaved re/im), run by cuBLAS through torch (library GEMM: the best case for
the tensor-core form, no fragment shuffles or TMEM round trips counted), in the precisions the tensor cores
offer, against (a) the accuracy the path needs (1e-5 relative power) and (b) the time of the WHOLE fused FP32
kernel (two radix-32 passes + decode + window + dB), 0.88 ms for the same points.
    python profiles/tc_ablation.py
"""
import json

import numpy as np
import torch

dev = "cuda:0"
M, R = 1 << 23, 32
k = np.arange(R)
W = np.exp(-2j * np.pi * np.outer(k, k) / R)                # x_row (1 x 32 complex) . W
B = np.zeros((2 * R, 2 * R))
B[0::2, 0::2], B[0::2, 1::2] = W.real, W.imag                 # [re im] . [[Wr Wi] [-Wi Wr]]
B[1::2, 0::2], B[1::2, 1::2] = -W.imag, W.real
g = torch.Generator(device=dev).manual_seed(1)
A32 = torch.randn(M, 2 * R, device=dev, generator=g, dtype=torch.float32)
B64 = torch.from_numpy(B).to(dev)
ref = (A32[:4096].double() @ B64)
ref_pow = ref[:, 0::2] ** 2 + ref[:, 1::2] ** 2


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]


def power_err(out):
    o = out[:4096].double()
    p = o[:, 0::2] ** 2 + o[:, 1::2] ** 2
    return float(((p - ref_pow).abs() / ref_pow.clamp_min(1e-30)).median()), float(((p - ref_pow).abs().max() / ref_pow.max()))


def split(x, dt, terms):
    parts, r = [], x.clone()
    for _ in range(terms):
        h = r.to(dt)
        parts.append(h)
        r = r - h.float()
    return parts


res = []
out = torch.empty(M, 2 * R, device=dev, dtype=torch.float32)
torch.backends.cuda.matmul.allow_tf32 = False
Bf = B64.float()
ms = timed(lambda: torch.matmul(A32, Bf, out=out))
res.append({"variant": "fp32 GEMM (no tensor cores)", "ms": round(ms, 3), "rel_power_err_median/max": power_err(out)})
torch.backends.cuda.matmul.allow_tf32 = True
ms = timed(lambda: torch.matmul(A32, Bf, out=out))
res.append({"variant": "tf32 x1", "ms": round(ms, 3), "rel_power_err_median/max": power_err(out)})
for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
    for terms in (1, 2, 3):
        a_parts = split(A32, dt, terms)
        b_parts = split(Bf, dt, terms)
        pairs = [(i, j) for i in range(terms) for j in range(terms) if i + j < terms]
        outs = [torch.empty(M, 2 * R, device=dev, dtype=torch.float32) for _ in pairs]

        def run():
            for n_, (i, j) in enumerate(pairs):
                outs[n_] = torch.mm(a_parts[i], b_parts[j], out_dtype=torch.float32)      # FP32 accumulate AND FP32 output
        ms = timed(run)
        acc = sum(o[:4096].double() for o in outs)
        res.append({"variant": "%s split x%d (%d GEMMs, products only; the split of the data and the FP32 sum are extra)" % (name, terms, len(pairs)),
                    "ms": round(ms, 3), "rel_power_err_median/max": power_err(acc)})
        del a_parts, outs
for r in res:
    print(json.dumps(r))
print(json.dumps({"note": "whole fused FP32 spectrogram kernel on the same 2^28 points (two radix-32 passes, decode, window, dB): 0.88 ms; "
                          "HBM floor of ONE unfused stage (2 GiB in + 2 GiB out): 0.66 ms"}))
