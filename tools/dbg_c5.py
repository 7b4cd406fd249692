import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectral_analyzer_b200 as sa
from spectral_analyzer_b200 import synth
from oracle import c_oracle as co
eng = sa.Engine(0)
nfft = 65536
if len(sys.argv) > 1 and sys.argv[1] == "big":
    blk = synth.recording(nfft, "cf64_le", seed=5)
    frames = 1024
    raw = torch.from_numpy(blk).cuda().repeat(frames)
    out = torch.empty(frames * nfft * 8, dtype=torch.uint8, device="cuda")
    p = eng.make_params("cf64_le", nfft, nfft, "hann", n_frames=frames, out="f64")
    eng.spectrogram_device(raw.data_ptr(), raw.numel(), p, out.data_ptr(), out.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("big done", torch.isfinite(out.view(torch.float64)).all().item())
    del raw, out
raw = synth.recording(nfft * 5, "cf64_le", seed=5)
for win, hop in (("rect", nfft), ("hann", nfft // 2)):
    ref = co.spectrogram(raw, "cf64_le", 0, nfft, hop, win, 5)
    got = eng.spectrogram(raw, "cf64_le", nfft, 5, hop=hop, window=win, precision="f64", out_kind="f64")
    print(win, "finite", np.isfinite(got).all(), "maxerr", np.nanmax(np.abs(got - ref)), "rows bad", [int(np.abs(got[i]-ref[i]).max() > 1e-6) for i in range(5)])
