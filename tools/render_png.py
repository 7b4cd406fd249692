#!/usr/bin/env python
"""Whole-recording (or windowed) spectrogram of a SigMF recording as a PNG, rendered on the GPU.

    python tools/render_png.py capture.sigmf-meta out.png [--nfft 1024] [--width 2048] [--height 1024]
                               [--reduce max|mean|nearest] [--colormap Heatmap|Grayscale] [--min-db -160] [--max-db -30]
                               [--start-sample 0] [--frames-per-column N]   (default: the whole recording in --width columns)

What the JavaFX view shows for one screen (MainController.updateDisplay + renderSpectrogram, :980-999, :1261-1291),
for any span of the file: the samples stream to the GPU once, only the width x height RGBA canvas comes back.
Raw / WAV files: create the metadata first with spectral_analyzer_b200.sigmf.NonconformingDatasetHelper.
"""
import argparse
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def png_bytes(rgba):
    """Minimal PNG encoder (8-bit RGBA, no interlace): uint8 [H, W, 4] -> bytes."""
    rgba = np.ascontiguousarray(rgba, np.uint8)
    h, w, c = rgba.shape
    if c != 4:
        raise ValueError("expected [H, W, 4]")

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgba.reshape(h, w * 4)], axis=1).tobytes()   # filter 0 per row
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def plan_view(total_samples, nfft, width, start_sample=0, frames_per_column=None):
    """(frames_per_column, columns that hold data): the whole remaining recording in `width` columns by default."""
    avail = max(0, total_samples - start_sample) // nfft
    if frames_per_column is None:
        frames_per_column = max(1, -(-avail // width))
    used = min(width, -(-avail // frames_per_column)) if avail else 0
    return frames_per_column, used


def render(engine, helper, nfft, width, height, reduce="max", colormap="Heatmap", min_db=-160.0, max_db=-30.0,
           start_sample=0, frames_per_column=None, window="hann"):
    fpc, _ = plan_view(helper.total_samples, nfft, width, start_sample, frames_per_column)
    return engine.render_canvas(helper.getDataBuffer(), helper.datatype, nfft, width, height, helper.sample_rate,
                                window=window, start_sample=start_sample, frames_per_column=fpc, reduce=reduce,
                                colormap=colormap, min_db=min_db, max_db=max_db)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("meta")
    ap.add_argument("png")
    ap.add_argument("--nfft", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--reduce", default="max", choices=["nearest", "max", "mean"])
    ap.add_argument("--colormap", default="Heatmap", choices=["Heatmap", "Grayscale"])
    ap.add_argument("--window", default="hann")
    ap.add_argument("--min-db", type=float, default=-160.0)
    ap.add_argument("--max-db", type=float, default=-30.0)
    ap.add_argument("--start-sample", type=int, default=0)
    ap.add_argument("--frames-per-column", type=int, default=None)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    import spectral_analyzer_b200 as sa
    from spectral_analyzer_b200 import sigmf
    h = sigmf.SigMfHelper().load(args.meta)
    eng = sa.Engine(args.device)                      # raises without a B200: there is no CPU fallback
    try:
        px = render(eng, h, args.nfft, args.width, args.height, args.reduce, args.colormap, args.min_db, args.max_db,
                    args.start_sample, args.frames_per_column, args.window)
    finally:
        eng.close()
    with open(args.png, "wb") as f:
        f.write(png_bytes(px))
    fpc, used = plan_view(h.total_samples, args.nfft, args.width, args.start_sample, args.frames_per_column)
    print("%s: %d samples (%s), %d frames per column, %d of %d columns hold data -> %s" % (
        os.path.basename(args.meta), h.total_samples, h.datatype, fpc, used, args.width, args.png))


if __name__ == "__main__":
    main()
