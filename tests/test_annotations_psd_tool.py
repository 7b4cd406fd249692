"""tools/annotations_psd.py: annotation -> downconvert parameter derivation and batching on CPU with a stub engine;
the GPU test finds the tone of a synthetic annotated recording at its annotated frequency."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import annotations_psd as ap                              # noqa: E402
from spectral_analyzer_b200 import sigmf                  # noqa: E402

FS, FC, N = 1.0e6, 100.0e6, 1 << 18


def write_recording(tmp_path):
    """cf32 recording: tone at +125 kHz (samples 0..2^17) and at -200 kHz (2^17..2^18), weak noise; two annotations
    around them, one without band edges (skipped) and one too short for any PSD (skipped)."""
    rng = np.random.default_rng(3)
    t = np.arange(N)
    f = np.where(t < N // 2, 125e3, -200e3)
    x = 0.5 * np.exp(2j * np.pi * np.cumsum(f) / FS) + 0.01 * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    iq = np.empty(2 * N, np.float32)
    iq[0::2], iq[1::2] = x.real, x.imag
    (tmp_path / "r.sigmf-data").write_bytes(iq.tobytes())
    anns = [
        {"core:sample_start": 1000, "core:sample_count": 100000, "core:freq_lower_edge": FC + 125e3 - 20e3,
         "core:freq_upper_edge": FC + 125e3 + 20e3, "core:label": "tone A"},
        {"core:sample_start": N // 2 + 500, "core:sample_count": 60000, "core:freq_lower_edge": FC - 200e3 - 50e3,
         "core:freq_upper_edge": FC - 200e3 + 50e3, "core:label": "tone B"},
        {"core:sample_start": 10, "core:sample_count": 5000, "core:label": "no band"},
        {"core:sample_start": 10, "core:sample_count": 300, "core:freq_lower_edge": FC - 1e3, "core:freq_upper_edge": FC + 1e3},
    ]
    meta = {"global": {"core:datatype": "cf32_le", "core:sample_rate": FS, "core:version": "1.0.0"},
            "captures": [{"core:sample_start": 0, "core:frequency": FC}], "annotations": anns}
    (tmp_path / "r.sigmf-meta").write_text(json.dumps(meta))
    return tmp_path / "r.sigmf-meta"


def test_rows_follow_the_reference_derivation(tmp_path):
    h = sigmf.SigMfHelper().load(write_recording(tmp_path))
    rows, info = ap.annotation_rows(h)
    assert [i[0] for i in info] == ["tone A", "tone B"]
    (s0, c0, f0, d0, fast0), (s1, c1, f1, d1, _) = rows
    assert (s0, c0, d0, fast0) == (1000, 100000, 25, False)              # down = floor(fs / bw) = floor(1e6 / 40e3)
    assert abs(f0 - 0.125) < 1e-12                                       # (centre - fc) / fs, AnnotationController.java:329-335
    assert (s1, c1, d1) == (N // 2 + 500, 60000, 10) and abs(f1 + 0.2) < 1e-12
    assert ap.psd_size(100000 // 25, 8192) == 2048 and ap.psd_size(6000, 8192) == 4096 and ap.psd_size(9000, 8192) == 8192


class StubEngine:
    def __init__(self):
        self.calls = []

    def downconvert_psd_batch(self, buffer, datatype, sample_rate, annotations, psd_nfft=8192, want_iq=True, **kw):
        self.calls.append((len(buffer), datatype, sample_rate, list(annotations), psd_nfft, want_iq))
        return None, np.tile(np.arange(psd_nfft, dtype=np.float64), (len(annotations), 1))


def test_batches_by_psd_size_with_stub_engine(tmp_path):
    h = sigmf.SigMfHelper().load(write_recording(tmp_path))
    eng = StubEngine()
    out, rows, info = ap.run(eng, h, psd_nfft=8192)
    assert list(out["psd_nfft"]) == [2048, 4096] and len(eng.calls) == 2
    assert all(c[5] is False and c[1] == "cf32_le" and c[2] == FS for c in eng.calls)
    assert out["psd_db_0"].shape == (2048,) and out["psd_db_1"].shape == (4096,)
    f0 = out["freq_hz_0"]
    assert f0[1024] == FC + 125e3 and abs((f0[1] - f0[0]) - FS / 25 / 2048) < 1e-9     # axis centred on the annotation


@pytest.mark.gpu
def test_tool_end_to_end_finds_the_tones(tmp_path):
    meta = write_recording(tmp_path)
    out_path = tmp_path / "o.npz"
    ap.main([str(meta), str(out_path)])
    d = np.load(out_path)
    for i, f_tone in ((0, FC + 125e3), (1, FC - 200e3)):
        psd, f = d["psd_db_%d" % i], d["freq_hz_%d" % i]
        assert abs(f[int(np.argmax(psd))] - f_tone) <= 2 * (f[1] - f[0])
        assert psd.max() - np.median(psd) > 40
