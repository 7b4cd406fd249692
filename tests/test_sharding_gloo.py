"""world_size-2 (and 3) gloo runs of the multi-GPU host logic on CPU: frame blocks + halo + display
assembly must reproduce the unsharded image.  The per-rank compute here is the CPU oracle (the
checker); on GPUs the same functions wrap Engine.spectrogram (bench.py, test_gpu_multi)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spectral_analyzer_b200 import sharding, synth      # noqa: E402


def test_frame_blocks_cover_exactly():
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [sharding.frame_block(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_sample_span_has_halo():
    # 1024-pt, hop 512: block [10, 20) reads 10*512 .. 19*512+1024, i.e. a 512-sample halo past 20*512
    assert sharding.sample_span(0, 10, 20, 512, 1024) == (5120, 19 * 512 + 1024)
    assert sharding.sample_span(7, 3, 3, 512, 1024) == (7 + 3 * 512, 7 + 3 * 512)
    assert sharding.sample_span(0, 0, 4, 1024, 1024) == (0, 4096)          # reference framing: no halo


def test_annotation_shares_balanced():
    counts = [100, 90, 80, 10, 10, 10, 5, 5]
    shares = sharding.annotation_shares(counts, 2)
    assert sorted(sum(shares, [])) == list(range(8))
    loads = [sum(counts[i] for i in s) for s in shares]
    assert max(loads) <= (4 * sum(counts)) // (3 * 2) + 1          # LPT bound: 4/3 of the ideal makespan


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, nfft, hop, q):
    from oracle import c_oracle as co
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw = synth.recording((n_frames - 1) * hop + nfft - 100, "ci16_le", seed=9)      # last frame past EOF

    def compute(first_sample, nf):
        # each rank touches only its own sample span (+ halo): slice the buffer to prove it
        s0, s1 = sharding.sample_span(0, first_sample // hop, first_sample // hop + nf, hop, nfft)
        s1 = min(s1, raw.size // 4)
        local = raw[s0 * 4: s1 * 4]
        return torch.from_numpy(co.spectrogram(local, "ci16_le", 0, nfft, hop, "hann", nf, nthreads=1))

    f0, f1, rows = sharding.local_spectrogram(compute, n_frames, world, rank, 0, hop, nfft)
    full = sharding.gather_rows(rows, n_frames, dst=0)
    if rank == 0:
        ref = co.spectrogram(raw, "ci16_le", 0, nfft, hop, "hann", n_frames, nthreads=1)
        q.put((bool(np.array_equal(full.numpy(), ref)), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 37), (3, 10), (2, 1)])
def test_sharded_image_equals_unsharded(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, 256, 128, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, shape = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and shape == (n_frames, 256)


def _canvas_worker(rank, world, port, W, H, fpc, q):
    """Each rank renders its own canvas columns from its own samples (the checker stands in for
    Engine.render_canvas), then only canvas-sized tiles are gathered."""
    from oracle import c_oracle as co
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nfft, fs = 128, 1.0e6
    raw = synth.recording(W * fpc * nfft, "cu8", seed=13)
    c0, c1, f0, f1 = sharding.canvas_columns(W, fpc, world, rank)
    s0, s1 = sharding.sample_span(0, f0, f1, nfft, nfft)
    db = co.spectrogram(raw[s0 * 2: s1 * 2], "cu8", 0, nfft, nfft, "rect", f1 - f0, nthreads=1)
    tile = co.render_canvas(db[::fpc], H, fs, -120.0, -20.0, "Heatmap")            # [H, c1 - c0, 4], nearest frame
    full = sharding.gather_canvas(torch.from_numpy(tile), W, dst=0)
    if rank == 0:
        db_all = co.spectrogram(raw, "cu8", 0, nfft, nfft, "rect", W * fpc, nthreads=1)
        ref = co.render_canvas(db_all[::fpc], H, fs, -120.0, -20.0, "Heatmap")
        q.put((bool(np.array_equal(full.numpy(), ref)), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W", [(2, 9), (3, 4)])
def test_sharded_canvas_equals_unsharded(world, W):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_canvas_worker, args=(r, world, port, W, 40, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, shape = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and shape == (40, W, 4)
