"""Time-sharding of a long recording over N GPUs (one process per GPU, torch.distributed).

Frames are independent, so the spectrogram shards without any collective on the data path
(SURVEY.md 8e): rank r of W owns the contiguous frame block [f0, f1) and reads the samples
[start + f0*hop, start + (f1-1)*hop + nfft) -- its own time block plus an (nfft - hop)-sample halo on
the right -- straight from the host mapping.  Only display assembly crosses GPUs: `gather_rows`
collects the row blocks on one rank (NCCL over NVLink for CUDA tensors, gloo for CPU tensors).
The reference has no counterpart: it shows canvasW frames at a time (MainController.java:980-999).
"""
import numpy as np


def frame_block(n_frames, world, rank):
    """Contiguous, balanced frame range of `rank`: sizes differ by at most one frame."""
    base, extra = divmod(n_frames, world)
    f0 = rank * base + min(rank, extra)
    return f0, f0 + base + (1 if rank < extra else 0)


def sample_span(start_sample, f0, f1, hop, nfft):
    """Samples [s0, s1) the frame block [f0, f1) reads, halo included; (s0, s0) if the block is empty."""
    s0 = start_sample + f0 * hop
    if f1 <= f0:
        return s0, s0
    return s0, start_sample + (f1 - 1) * hop + nfft


def local_spectrogram(compute, n_frames, world, rank, start_sample, hop, nfft):
    """Runs `compute(first_frame_sample, n_local_frames)` for this rank's block and returns
    (f0, f1, rows).  `compute` is Engine.spectrogram bound to the recording (product path)."""
    f0, f1 = frame_block(n_frames, world, rank)
    rows = compute(start_sample + f0 * hop, f1 - f0)
    return f0, f1, rows


def gather_rows(rows, n_frames, group=None, dst=0):
    """Display assembly: concatenates the per-rank row blocks (torch tensors, [n_local, nfft...]) on
    rank `dst` in frame order; other ranks get None.  Uses all_gather on equal-sized padded blocks
    (one collective, NCCL- and gloo-compatible)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_frames + world - 1) // world
    block = torch.zeros((per,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    block[: rows.shape[0]] = rows
    out = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, out, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        f0, f1 = frame_block(n_frames, world, r)
        parts.append(out[r][: f1 - f0])
    return torch.cat(parts, dim=0)


def canvas_columns(canvas_w, frames_per_column, world, rank):
    """Canvas columns [c0, c1) of `rank` and the frame block they are made of: a rank renders whole
    columns (Engine.render_canvas on its own samples), so only canvas-sized tiles are gathered."""
    c0, c1 = frame_block(canvas_w, world, rank)
    return c0, c1, c0 * frames_per_column, c1 * frames_per_column


def gather_canvas(tile, canvas_w, group=None, dst=0):
    """Display assembly of the decimated image (gather the display-sized image, not the
    full-resolution rows): per-rank tiles [canvas_h, c1 - c0, 4] uint8 -> [canvas_h, canvas_w, 4] on `dst`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (canvas_w + world - 1) // world
    h = tile.shape[0]
    block = torch.zeros((per, h) + tuple(tile.shape[2:]), dtype=tile.dtype, device=tile.device)
    block[: tile.shape[1]] = tile.transpose(0, 1)              # column-major blocks concatenate along columns
    out = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, out, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        c0, c1 = frame_block(canvas_w, world, r)
        parts.append(out[r][: c1 - c0])
    return torch.cat(parts, dim=0).transpose(0, 1).contiguous()


def annotation_shares(counts, world):
    """Size-balanced assignment of annotations to ranks (longest-processing-time first).
    Returns a list of index lists, one per rank."""
    order = np.argsort(-np.asarray(counts, dtype=np.int64), kind="stable")
    load = [0] * world
    shares = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        shares[r].append(int(i))
        load[r] += int(counts[i])
    return [sorted(s) for s in shares]
