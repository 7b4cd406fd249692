import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import spectral_analyzer_b200 as sa
n = 1 << 28; nfft, hop = 1024, 512
frames = (n - nfft) // hop + 1
raw = np.random.default_rng(0).standard_normal(2 * n, dtype=np.float32)
path = "/dev/shm/mm_test.bin"; raw.tofile(path)
mm = np.memmap(path, dtype=np.uint8, mode="r")
out = np.empty((frames, nfft), np.float32)
pin_out = torch.empty((frames, nfft), dtype=torch.float32, pin_memory=True).numpy()
eng = sa.Engine(0)
def t(fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
print("threads", os.environ.get("SA_COPY_THREADS"), "chunk", os.environ.get("SA_CHUNK_MB"),
      "mmap->pageable %.1f ms" % t(lambda: eng.spectrogram(mm, "cf32_le", nfft, frames, hop=hop, window="hann", out=out)),
      "mmap->pinned %.1f ms" % t(lambda: eng.spectrogram(mm, "cf32_le", nfft, frames, hop=hop, window="hann", out=pin_out)),
      "file->pinned %.1f ms" % t(lambda: eng.spectrogram_file(path, "cf32_le", nfft, frames, hop=hop, window="hann", out=pin_out)))
os.unlink(path)
