"""Property tests (hypothesis) of the host-side index logic: frame / column partitioning with halos, the canvas tile
cache's composition for arbitrary scroll positions, and multi-capture byte spans.  Index work must be exact."""
import json

import numpy as np
from hypothesis import given, settings, strategies as st

from spectral_analyzer_b200 import sharding, sigmf
from spectral_analyzer_b200.tiles import CanvasTileCache


@settings(max_examples=300, deadline=None)
@given(n_frames=st.integers(0, 10 ** 7), world=st.integers(1, 16), hop=st.integers(1, 1 << 17),
       log2n=st.integers(6, 16), start=st.integers(0, 1 << 40))
def test_frame_blocks_partition_and_spans_carry_the_halo(n_frames, world, hop, log2n, start):
    nfft = 1 << log2n
    prev = 0
    sizes = []
    for r in range(world):
        f0, f1 = sharding.frame_block(n_frames, world, r)
        assert f0 == prev and f1 >= f0
        prev = f1
        sizes.append(f1 - f0)
        s0, s1 = sharding.sample_span(start, f0, f1, hop, nfft)
        if f1 > f0:
            # exactly the samples of frames f0 .. f1-1: first frame's start to last frame's end
            assert s0 == start + f0 * hop and s1 == start + (f1 - 1) * hop + nfft
            if r + 1 < world and sizes[-1] and hop < nfft:
                nxt0, _ = sharding.sample_span(start, f1, f1 + 1, hop, nfft)
                assert s1 - nxt0 == nfft - hop                      # halo shared with the next rank's first frame
        else:
            assert s0 == s1
    assert prev == n_frames and max(sizes) - min(sizes) <= 1


@settings(max_examples=200, deadline=None)
@given(canvas_w=st.integers(1, 5000), fpc=st.integers(1, 1000), world=st.integers(1, 8))
def test_canvas_columns_partition(canvas_w, fpc, world):
    prev_c, prev_f = 0, 0
    for r in range(world):
        c0, c1, f0, f1 = sharding.canvas_columns(canvas_w, fpc, world, r)
        assert (c0, f0) == (prev_c, prev_f) and f1 - f0 == (c1 - c0) * fpc
        prev_c, prev_f = c1, f1
    assert prev_c == canvas_w and prev_f == canvas_w * fpc


@settings(max_examples=200, deadline=None)
@given(counts=st.lists(st.integers(1, 1 << 24), min_size=0, max_size=60), world=st.integers(1, 8))
def test_annotation_shares_are_a_partition_and_lpt_balanced(counts, world):
    shares = sharding.annotation_shares(counts, world)
    flat = sorted(i for s in shares for i in s)
    assert flat == list(range(len(counts))) and len(shares) == world
    if counts:
        loads = [sum(counts[i] for i in s) for s in shares]
        assert max(loads) - min(loads) <= max(counts)             # longest-processing-time-first bound


class StubEngine:
    def __init__(self):
        self.calls = 0

    def render_canvas(self, buffer, datatype, nfft, canvas_w, canvas_h, sample_rate, hop=None, window="rect",
                      start_sample=0, frames_per_column=1, reduce="nearest", **kw):
        self.calls += 1
        spc = frames_per_column * (hop or nfft)
        col = (start_sample + spc * np.arange(canvas_w, dtype=np.uint64)).astype(np.uint32)
        return np.broadcast_to(col.view(np.uint8).reshape(1, canvas_w, 4), (canvas_h, canvas_w, 4)).copy()


@settings(max_examples=150, deadline=None)
@given(starts=st.lists(st.integers(0, 1 << 22), min_size=1, max_size=6), w=st.integers(1, 300),
       tile_w=st.integers(1, 128), fpc=st.integers(1, 5), snap=st.booleans(), max_tiles=st.integers(1, 6))
def test_tile_cache_view_is_exact_for_any_scroll_sequence(starts, w, tile_w, fpc, snap, max_tiles):
    eng = StubEngine()
    cache = CanvasTileCache(eng, tile_w=tile_w, max_tiles=max_tiles)
    buf = np.zeros(8, np.uint8)
    nfft = 64
    spc = nfft * fpc
    for s in starts:
        v = cache.view(buf, "cu8", nfft, w, 2, 1.0, start_sample=s, frames_per_column=fpc, snap=snap)
        first = s - s % spc if snap else s
        got = v[0].copy().view(np.uint32).reshape(-1)
        assert np.array_equal(got, (first + spc * np.arange(w)).astype(np.uint32))
        assert len(cache) <= max_tiles
    assert cache.hits + cache.misses == eng.calls + cache.hits


@settings(max_examples=100, deadline=None)
@given(lens=st.lists(st.integers(0, 500), min_size=1, max_size=5), hdrs=st.lists(st.integers(0, 64), min_size=5, max_size=5),
       dt=st.sampled_from(["cu8", "ci16_le", "cf32_le", "cf64_le"]))
def test_capture_segments_tile_the_file(tmp_path_factory, lens, hdrs, dt):
    bps = sigmf.bytes_per_sample(dt)
    d = tmp_path_factory.mktemp("mc")
    rng = np.random.default_rng(sum(lens) + len(lens))
    parts = [rng.integers(0, 256, n * bps, dtype=np.uint8).tobytes() for n in lens]
    heads = [bytes([65 + i]) * hdrs[i] for i in range(len(lens))]
    (d / "x.bin").write_bytes(b"".join(h + p for h, p in zip(heads, parts)))
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
    meta = {"global": {"core:datatype": dt, "core:sample_rate": 1.0, "core:version": "1.0.0", "core:dataset": "x.bin"},
            "captures": [{"core:sample_start": int(s), "core:header_bytes": len(h)} for s, h in zip(starts, heads)],
            "annotations": []}
    (d / "x.sigmf-meta").write_text(json.dumps(meta))
    h = sigmf.SigMfHelper().load(d / "x.sigmf-meta")
    segs = h.capture_segments()
    assert [s[1] for s in segs] == starts and [s[2] for s in segs] == lens
    for i, p in enumerate(parts):
        assert bytes(h.capture_buffer(i)) == p
