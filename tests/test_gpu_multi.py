"""Multi-GPU path on real devices (skipped on a one-GPU box): one process per GPU over NCCL, every rank computes
its time block (+ halo) with its own Engine, display assembly through sharding.gather_rows / gather_canvas; the
result must equal the one-GPU image bit for bit (frames are independent: no collective on the data path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import spectral_analyzer_b200 as sa
    from spectral_analyzer_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nfft, hop, n_frames = 1024, 512, 1001
    raw = synth.recording((n_frames - 1) * hop + nfft - 300, "ci16_le", seed=17)        # last frame past EOF
    eng = sa.Engine(rank)

    def compute(first_sample, nf):
        return torch.from_numpy(eng.spectrogram(raw, "ci16_le", nfft, nf, hop=hop, window="hann",
                                                start_sample=first_sample)).cuda()
    f0, f1, rows = sharding.local_spectrogram(compute, n_frames, world, rank, 0, hop, nfft)
    full = sharding.gather_rows(rows, n_frames, dst=0)
    # display-decimated assembly: every rank renders whole canvas columns from its own samples
    W, H, fpc = 90, 128, 11
    c0, c1, cf0, cf1 = sharding.canvas_columns(W, fpc, world, rank)
    tile = eng.render_canvas(raw, "ci16_le", nfft, c1 - c0, H, 2.0e6, hop=hop, window="hann", start_sample=cf0 * hop,
                             frames_per_column=fpc, reduce="max", colormap="Heatmap")
    canvas = sharding.gather_canvas(torch.from_numpy(tile).cuda(), W, dst=0)
    if rank == 0:
        ref = eng.spectrogram(raw, "ci16_le", nfft, n_frames, hop=hop, window="hann")
        ref_canvas = eng.render_canvas(raw, "ci16_le", nfft, W, H, 2.0e6, hop=hop, window="hann", frames_per_column=fpc,
                                       reduce="max", colormap="Heatmap")
        q.put((bool(np.array_equal(full.cpu().numpy(), ref)), bool(np.array_equal(canvas.cpu().numpy(), ref_canvas)),
               tuple(full.shape), tuple(canvas.shape)))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def test_two_gpus_reproduce_the_one_gpu_image():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows_ok, canvas_ok, shape, cshape = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert rows_ok and canvas_ok and shape == (1001, 1024) and cshape == (128, 90, 4)
