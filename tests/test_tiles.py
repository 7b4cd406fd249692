"""Canvas tile cache: composition and LRU logic against a stub renderer on CPU; against
Engine.render_canvas on the GPU (bit-identical to one direct render at the same view start)."""
import numpy as np
import pytest

from spectral_analyzer_b200.tiles import CanvasTileCache


class StubEngine:
    """Every pixel of a column carries the column's first sample index (mod 2^32), little endian."""

    def __init__(self):
        self.calls = []

    def render_canvas(self, buffer, datatype, nfft, canvas_w, canvas_h, sample_rate, hop=None, window="rect",
                      start_sample=0, frames_per_column=1, reduce="nearest", **kw):
        self.calls.append((start_sample, canvas_w))
        spc = frames_per_column * (hop or nfft)
        col = (start_sample + spc * np.arange(canvas_w, dtype=np.uint64)).astype(np.uint32)
        return np.broadcast_to(col.view(np.uint8).reshape(1, canvas_w, 4), (canvas_h, canvas_w, 4)).copy()


def column_starts(view):
    return view[0].copy().view(np.uint32).reshape(-1)


def test_view_composition_and_reuse():
    eng = StubEngine()
    cache = CanvasTileCache(eng, tile_w=64, max_tiles=16)
    buf = np.zeros(16, np.uint8)
    nfft, W, H, fpc = 256, 200, 8, 3
    spc = nfft * fpc
    for start in (0, 5 * spc, 5 * spc + 17, 70 * spc + 17, 64 * spc):
        v = cache.view(buf, "cf32_le", nfft, W, H, 1e6, start_sample=start, frames_per_column=fpc, reduce="max")
        assert v.shape == (H, W, 4)
        assert np.array_equal(column_starts(v), (start + spc * np.arange(W)).astype(np.uint32))
        for s, w in eng.calls:                               # tiles start on tile boundaries of the view's phase
            assert w == 64 and (s - s % spc) // spc % 64 == 0
    # views 0 and 5*spc share phase 0: the second needed no new tile beyond the first's 4 (columns 0..255)
    assert eng.calls[:4] == [(k * 64 * spc, 64) for k in range(4)]
    assert cache.hits > 0


def test_scroll_renders_only_entering_tiles_and_lru_evicts():
    eng = StubEngine()
    cache = CanvasTileCache(eng, tile_w=32, max_tiles=5)
    buf = np.zeros(16, np.uint8)
    nfft, W = 64, 128
    cache.view(buf, "cu8", nfft, W, 4, 1.0, start_sample=0)
    assert len(eng.calls) == 4
    cache.view(buf, "cu8", nfft, W, 4, 1.0, start_sample=32 * nfft)          # one tile enters
    assert len(eng.calls) == 5 and len(cache) == 5
    cache.view(buf, "cu8", nfft, W, 4, 1.0, start_sample=64 * nfft)          # one more: the oldest is evicted
    assert len(eng.calls) == 6 and len(cache) == 5
    cache.view(buf, "cu8", nfft, W, 4, 1.0, start_sample=0)                  # tile 0 was evicted, 1..3: 1 evicted too by now
    assert len(eng.calls) > 6
    # a different rendering parameter never reuses a tile
    n = len(eng.calls)
    cache.view(buf, "cu8", nfft, W, 4, 1.0, start_sample=0, colormap="Heatmap")
    assert len(eng.calls) == n + 4


def test_snap_shares_tiles_between_phases():
    eng = StubEngine()
    cache = CanvasTileCache(eng, tile_w=64)
    buf = np.zeros(16, np.uint8)
    nfft, W = 128, 64
    a = cache.view(buf, "ci16_le", nfft, W, 2, 1.0, start_sample=10 * nfft + 5, snap=True)
    n = len(eng.calls)
    b = cache.view(buf, "ci16_le", nfft, W, 2, 1.0, start_sample=10 * nfft + 99, snap=True)
    assert len(eng.calls) == n and np.array_equal(a, b)
    assert column_starts(a)[0] == 10 * nfft
    with pytest.raises(ValueError):
        cache.view(buf, "ci16_le", nfft, W, 2, 1.0, start_sample=-1)


@pytest.mark.gpu
@pytest.mark.parametrize("reduce,fpc", [("nearest", 1), ("max", 4), ("mean", 3)])
def test_gpu_tiled_view_equals_direct_render(engine, reduce, fpc):
    from spectral_analyzer_b200 import synth
    nfft, W, H = 256, 300, 128
    n = nfft * fpc * 700 + 77
    raw = synth.recording(n, "ci16_le", seed=9)
    cache = CanvasTileCache(engine, tile_w=128, max_tiles=8)
    kw = dict(hop=nfft, window="hann", frames_per_column=fpc, reduce=reduce, colormap="Heatmap")
    # the third view runs past EOF; the fourth overlaps tiles of the second and third (same phase)
    for start in (0, 3 * nfft * fpc + 11, 450 * nfft * fpc + 11, 200 * nfft * fpc + 11):
        direct = engine.render_canvas(raw, "ci16_le", nfft, W, H, 1e6, start_sample=start, **kw)
        tiled = cache.view(raw, "ci16_le", nfft, W, H, 1e6, start_sample=start, **kw)
        assert np.array_equal(direct, tiled)
    assert cache.hits > 0
