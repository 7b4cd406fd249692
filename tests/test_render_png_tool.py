"""tools/render_png.py: the host-side pieces (PNG encoder, view planning, SigMF glue) on CPU with a stub engine; the
GPU test renders a synthetic recording end to end."""
import json
import os
import struct
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import render_png as rp                                   # noqa: E402
from spectral_analyzer_b200 import sigmf, synth           # noqa: E402


def decode_png(b):
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(b):
        n = struct.unpack(">I", b[pos:pos + 4])[0]
        tag, data = b[pos + 4:pos + 8], b[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        chunks.append((tag, data))
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[0][1][:10])
    assert (depth, ctype) == (8, 6) and chunks[-1][0] == b"IEND"
    raw = np.frombuffer(zlib.decompress(b"".join(d for t, d in chunks if t == b"IDAT")), np.uint8).reshape(h, 1 + 4 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 4)


def test_png_roundtrip():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    assert np.array_equal(decode_png(rp.png_bytes(img)), img)
    with pytest.raises(ValueError):
        rp.png_bytes(np.zeros((4, 4, 3), np.uint8))


def test_plan_view():
    assert rp.plan_view(1 << 20, 1024, 256) == (4, 256)                 # 1024 frames in 256 columns
    assert rp.plan_view(1000 * 1024 + 5, 1024, 256) == (4, 250)          # ragged tail: 250 columns hold data
    assert rp.plan_view(100 * 1024, 1024, 256) == (1, 100)               # short recording: one frame per column
    assert rp.plan_view(1 << 20, 1024, 256, start_sample=1 << 19) == (2, 256)
    assert rp.plan_view(1 << 20, 1024, 256, frames_per_column=1) == (1, 256)
    assert rp.plan_view(0, 1024, 256) == (1, 0)


class StubEngine:
    def render_canvas(self, buffer, datatype, nfft, canvas_w, canvas_h, sample_rate, **kw):
        self.call = dict(n=len(buffer), datatype=datatype, nfft=nfft, w=canvas_w, h=canvas_h, fs=sample_rate, **kw)
        return np.full((canvas_h, canvas_w, 4), 255, np.uint8)


def write_recording(tmp_path, n=64 * 1024):
    raw = synth.recording(n, "ci16_le", seed=4).tobytes()
    (tmp_path / "a.sigmf-data").write_bytes(raw)
    meta = {"global": {"core:datatype": "ci16_le", "core:sample_rate": 2.0e6, "core:version": "1.0.0"},
            "captures": [{"core:sample_start": 0, "core:frequency": 1e8}], "annotations": []}
    (tmp_path / "a.sigmf-meta").write_text(json.dumps(meta))
    return tmp_path / "a.sigmf-meta"


def test_render_glue_with_stub_engine(tmp_path):
    h = sigmf.SigMfHelper().load(write_recording(tmp_path))
    eng = StubEngine()
    px = rp.render(eng, h, 1024, 16, 8, reduce="mean", colormap="Grayscale")
    assert px.shape == (8, 16, 4)
    c = eng.call
    assert (c["n"], c["datatype"], c["nfft"], c["fs"]) == (64 * 1024 * 4, "ci16_le", 1024, 2.0e6)
    assert c["frames_per_column"] == 4 and c["reduce"] == "mean" and c["colormap"] == "Grayscale" and c["start_sample"] == 0


@pytest.mark.gpu
def test_tool_end_to_end(tmp_path):
    meta = write_recording(tmp_path)
    out = tmp_path / "a.png"
    rp.main([str(meta), str(out), "--nfft", "256", "--width", "64", "--height", "32"])
    img = decode_png(out.read_bytes())
    assert img.shape == (32, 64, 4) and (img[..., 3] == 255).all() and img[..., :3].max() > 0
