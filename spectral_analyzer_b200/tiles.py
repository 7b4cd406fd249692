"""Canvas tile cache for scrolling.

The reference redraws the whole canvas on every scroll-bar move: canvasW calls of computeMagnitudes plus one
renderSpectrogram (S/controllers/MainController.java:980-999, :1261-1291; the offset comes from the scroll bar at
:319).  Here a view is assembled from fixed-width tiles of canvas columns rendered by Engine.render_canvas
(sa_render_canvas) and kept in an LRU map, so that a scroll step only renders the columns that enter the view.

A canvas column c of a view that starts at sample `start` covers samples
[start + c * spc, start + (c + 1) * spc), spc = frames_per_column * hop.  Tiles are aligned on GLOBAL column
indices g = (start - phase) / spc with phase = start % spc, so two views share tiles iff they have the same phase;
`snap=True` rounds the view start down to a column boundary (phase 0) so that every scroll position shares one set
of tiles, at the price of moving the view by less than one column.  Without snapping the result is bit-identical
to one render_canvas call at the same start (each column's pixels depend on that column's samples only).
"""
import collections

import numpy as np


class CanvasTileCache:
    def __init__(self, engine, tile_w=256, max_tiles=64):
        if tile_w < 1 or max_tiles < 1:
            raise ValueError("tile_w and max_tiles must be positive")
        self.engine, self.tile_w, self.max_tiles = engine, int(tile_w), int(max_tiles)
        self._tiles = collections.OrderedDict()
        self.hits = self.misses = 0

    def clear(self):
        self._tiles.clear()

    def __len__(self):
        return len(self._tiles)

    @staticmethod
    def _buffer_id(buffer):
        a = buffer if isinstance(buffer, np.ndarray) else np.frombuffer(buffer, dtype=np.uint8)
        return (a.__array_interface__["data"][0], a.nbytes)

    def view(self, buffer, datatype, nfft, canvas_w, canvas_h, sample_rate, start_sample=0, hop=None, window="rect",
             frames_per_column=1, reduce="nearest", colormap="Grayscale", min_db=-160.0, max_db=-30.0, snap=False,
             **kw):
        """uint8 [canvas_h, canvas_w, 4] for the view whose first column starts at `start_sample`
        (currentSampleOffset, MainController.java:984)."""
        if start_sample < 0:
            raise ValueError("start_sample must be non-negative")
        hop = nfft if hop is None else hop
        spc = frames_per_column * hop
        phase = 0 if snap else start_sample % spc
        g0 = (start_sample - (start_sample % spc)) // spc          # global index of the view's first column
        base = (self._buffer_id(buffer), datatype, nfft, hop, window, canvas_h, frames_per_column, reduce, colormap,
                float(min_db), float(max_db), float(sample_rate), phase, tuple(sorted(kw.items())))
        out = np.empty((canvas_h, canvas_w, 4), np.uint8)
        tw = self.tile_w
        for k in range(g0 // tw, (g0 + canvas_w - 1) // tw + 1):
            key = base + (k,)
            tile = self._tiles.get(key)
            if tile is None:
                self.misses += 1
                tile = self.engine.render_canvas(buffer, datatype, nfft, tw, canvas_h, sample_rate, hop=hop,
                                                 window=window, start_sample=phase + k * tw * spc,
                                                 frames_per_column=frames_per_column, reduce=reduce, colormap=colormap,
                                                 min_db=min_db, max_db=max_db, **kw)
                self._tiles[key] = tile
                while len(self._tiles) > self.max_tiles:
                    self._tiles.popitem(last=False)
            else:
                self.hits += 1
                self._tiles.move_to_end(key)
            lo, hi = max(g0, k * tw), min(g0 + canvas_w, (k + 1) * tw)
            out[:, lo - g0:hi - g0] = tile[:, lo - k * tw:hi - k * tw]
        return out
