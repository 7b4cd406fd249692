"""Multi-rank path on real devices: one process per rank, every rank computes its time block (+ halo) with its own
Engine, display assembly through sharding.gather_rows / gather_canvas; the result must equal the one-GPU image bit for
bit (frames are independent: no collective on the data path).  With two or more GPUs: one GPU per rank, NCCL.  On a
one-GPU box the same test still runs: two ranks share cuda:0 (NCCL refuses two ranks on one device, so the display
assembly goes over gloo on host tensors) -- the CUDA kernels, the sharding arithmetic and the halo reads are the same."""
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


def _worker(rank, world, port, q, one_gpu=False):
    import torch.distributed as dist
    import spectral_analyzer_b200 as sa
    from spectral_analyzer_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if one_gpu else rank
    torch.cuda.set_device(dev)
    if one_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    place = (lambda x: x) if one_gpu else (lambda x: x.cuda())
    nfft, hop, n_frames = 1024, 512, 1001
    raw = synth.recording((n_frames - 1) * hop + nfft - 300, "ci16_le", seed=17)        # last frame past EOF
    eng = sa.Engine(dev)

    def compute(first_sample, nf):
        return place(torch.from_numpy(eng.spectrogram(raw, "ci16_le", nfft, nf, hop=hop, window="hann",
                                                      start_sample=first_sample)))
    f0, f1, rows = sharding.local_spectrogram(compute, n_frames, world, rank, 0, hop, nfft)
    full = sharding.gather_rows(rows, n_frames, dst=0)
    # display-decimated assembly: every rank renders whole canvas columns from its own samples
    W, H, fpc = 90, 128, 11
    c0, c1, cf0, cf1 = sharding.canvas_columns(W, fpc, world, rank)
    tile = eng.render_canvas(raw, "ci16_le", nfft, c1 - c0, H, 2.0e6, hop=hop, window="hann", start_sample=cf0 * hop,
                             frames_per_column=fpc, reduce="max", colormap="Heatmap")
    canvas = sharding.gather_canvas(place(torch.from_numpy(tile)), W, dst=0)
    if rank == 0:
        ref = eng.spectrogram(raw, "ci16_le", nfft, n_frames, hop=hop, window="hann")
        ref_canvas = eng.render_canvas(raw, "ci16_le", nfft, W, H, 2.0e6, hop=hop, window="hann", frames_per_column=fpc,
                                       reduce="max", colormap="Heatmap")
        q.put((bool(np.array_equal(full.cpu().numpy(), ref)), bool(np.array_equal(canvas.cpu().numpy(), ref_canvas)),
               tuple(full.shape), tuple(canvas.shape)))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def _run(world, one_gpu):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, one_gpu)) for r in range(world)]
    for p in procs:
        p.start()
    rows_ok, canvas_ok, shape, cshape = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert rows_ok and canvas_ok and shape == (1001, 1024) and cshape == (128, 90, 4)


def test_two_gpus_reproduce_the_one_gpu_image():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    _run(world, one_gpu=False)


def test_three_ranks_sharing_one_gpu_reproduce_the_one_gpu_image():
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    _run(3, one_gpu=True)
