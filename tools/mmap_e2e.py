"""End-to-end spectrogram from an mmapped .sigmf-data FILE (what SigMfHelper hands the engine): pageable mapping vs
cudaHostRegister'ed mapping vs a pinned copy.   python tools/mmap_e2e.py [log2_samples]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectral_analyzer_b200 as sa                      # noqa: E402
from spectral_analyzer_b200 import synth                 # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n = 1 << log2n
path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "sa_mmap_test.sigmf-data")
blk = synth.recording(1 << 20, "cf32_le", seed=1)
with open(path, "wb") as f:
    for _ in range(n >> 20):
        f.write(blk.tobytes())
buf = np.memmap(path, dtype=np.uint8, mode="r")
eng = sa.Engine(0)
frames = (n - 1024) // 512 + 1
out = np.empty((frames, 1024), np.float32)


def run(tag, b):
    eng.spectrogram(b, "cf32_le", 1024, frames, hop=512, window="hann", out=out)
    t0 = time.perf_counter()
    for _ in range(3):
        eng.spectrogram(b, "cf32_le", 1024, frames, hop=512, window="hann", out=out)
    dt = (time.perf_counter() - t0) / 3
    print("%-34s %8.1f ms  %8.1f Msamples/s  in %.1f GB/s" % (tag, dt * 1e3, frames * 512 / dt / 1e6, n * 8 / dt / 1e9))


run("pageable mmap, pageable out", buf)
try:
    eng.register_host(buf, read_only=True)
    run("registered mmap, pageable out", buf)
    eng.register_host(out, read_only=False)
    run("registered mmap, registered out", buf)
    eng.unregister_host(out)
    eng.unregister_host(buf)
except sa.EngineError as e:
    print("register failed:", e)
import torch
pin_in = torch.empty(n * 8, dtype=torch.uint8, pin_memory=True)
pin_in.numpy()[:] = buf
pin_out = torch.empty((frames, 1024), dtype=torch.float32, pin_memory=True)
out = pin_out.numpy()
run("pinned copy in, pinned out", pin_in.numpy())
os.remove(path)
