"""The tensor-core ablation kernel (csrc/tc_ablation.cu: first radix-32 pass of the 1024-point spectrogram as a tcgen05
TF32x3 GEMM with TMEM accumulators) must meet the same parity tolerances as the shipped FP32 kernel -- north_star only
allows a tensor-core stage "within tolerance".  The timing comparison lives in tools/tc_ablation.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from spectral_analyzer_b200 import synth, _capi
from util import check_db_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("window,hop", [("hann", 512), ("rect", 1024), ("blackman_harris", 256)])
def test_tcgen05_pass_meets_the_fp32_tolerances(engine, window, hop):
    nfft, frames = 1024, 37
    raw = synth.recording((frames - 2) * hop + nfft + 8, "cf32_le", seed=71)      # the last frame runs past EOF
    d_iq = torch.from_numpy(raw.copy()).cuda()
    d_out = torch.zeros((frames, nfft), dtype=torch.float32, device="cuda")
    p = engine.make_params("cf32_le", nfft, hop, window, n_frames=frames, start_sample=8)
    _capi.check(_capi.lib().sa_ablation_tc_spectrogram_device(engine.handle, d_iq.data_ptr(), d_iq.numel(), C.byref(p),
                                                              d_out.data_ptr(), d_out.numel() * 4, 0,
                                                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert engine.last_kernel.startswith("tc_spectrogram_kernel")
    got = d_out.cpu().numpy()
    ref = co.spectrogram(raw, "cf32_le", 8, nfft, hop, window, frames)
    assert (ref[-1] == -150.0).all() and (got[-1] == -150.0).all()
    check_db_parity(got[:-1], ref[:-1])
    # the product path never takes this kernel
    engine.spectrogram(raw, "cf32_le", nfft, 3, hop=hop, window=window)
    assert not engine.last_kernel.startswith("tc_")
