"""SigMF ingest: host-side mirror of SigMfHelper / Global / NonconformingDatasetHelper and the
annotation parameter derivation.  CPU tests check the metadata rules; the GPU test feeds a WAV-derived mapping
(44-byte header, so the samples are only 4-byte aligned in the file) through the engine."""
import json
import os
import struct
import wave

import numpy as np
import pytest

from oracle import c_oracle as co
from spectral_analyzer_b200 import sigmf, synth
from util import check_db_parity


def write_sigmf(tmp_path, name, datatype, payload, header=b"", dataset=None, fs=2.0e6, fc=100e6, annotations=()):
    meta = {"global": {"core:datatype": datatype, "core:sample_rate": fs, "core:version": "1.0.0"},
            "captures": [{"core:sample_start": 0, "core:frequency": fc}], "annotations": list(annotations)}
    if header:
        meta["captures"][0]["core:header_bytes"] = len(header)
    data_name = dataset or (name + ".sigmf-data")
    if dataset:
        meta["global"]["core:dataset"] = dataset
    (tmp_path / data_name).write_bytes(header + payload)
    mp = tmp_path / (name + ".sigmf-meta")
    mp.write_text(json.dumps(meta))
    return mp


@pytest.mark.parametrize("dt,bps", [("cf32_le", 8), ("cf32_be", 8), ("ci16_le", 4), ("cu8", 2), ("ci8", 2), ("cf64_le", 16),
                                    ("ri16_le", 8), ("bogus", 8)])
def test_bytes_per_sample_matches_global_java(dt, bps):
    assert sigmf.bytes_per_sample(dt) == bps            # Global.java:67-79 incl. the fallback


def test_load_conforming_pair_and_header_bytes(tmp_path):
    raw = synth.recording(5000, "ci16_le", seed=3).tobytes()
    mp = write_sigmf(tmp_path, "a", "ci16_le", raw)
    h = sigmf.SigMfHelper().load(mp)
    assert bytes(h.getDataBuffer()) == raw and h.total_samples == 5000
    assert h.datatype == "ci16_le" and h.sample_rate == 2.0e6 and h.center_frequency == 100e6
    # header bytes of the first capture are skipped (SigMfHelper.java:59-67, :84)
    mp = write_sigmf(tmp_path, "b", "ci16_le", raw, header=b"H" * 44)
    h = sigmf.SigMfHelper().load(mp)
    assert bytes(h.getDataBuffer()) == raw
    # non-conforming dataset named in the global section (:49-52)
    mp = write_sigmf(tmp_path, "c", "cu8", raw, dataset="capture.cu8")
    h = sigmf.SigMfHelper().load(mp)
    assert bytes(h.getDataBuffer()) == raw and h.dataPath.endswith("capture.cu8")
    # header longer than the file -> empty buffer (Math.max(0, ...), :76)
    mp = write_sigmf(tmp_path, "d", "cu8", b"", header=b"")
    meta = json.loads(mp.read_text()); meta["captures"][0]["core:header_bytes"] = 10; mp.write_text(json.dumps(meta))
    assert len(sigmf.SigMfHelper().load(mp).getDataBuffer()) == 0


def test_missing_data_file_raises(tmp_path):
    mp = tmp_path / "x.sigmf-meta"
    mp.write_text(json.dumps({"global": {"core:datatype": "cf32_le"}, "captures": [], "annotations": []}))
    with pytest.raises(OSError):
        sigmf.SigMfHelper().load(mp)


def test_save_roundtrip_keeps_global_and_captures(tmp_path):
    mp = write_sigmf(tmp_path, "a", "cf32_le", b"\0" * 64)
    h = sigmf.SigMfHelper().load(mp)
    ann = [{"core:sample_start": 2, "core:sample_count": 4, "core:freq_lower_edge": 1.0, "core:freq_upper_edge": 2.0,
            "core:label": "x"}]
    h.saveSigMF(ann)
    h2 = sigmf.SigMfHelper().load(mp)
    assert h2.getParsedAnnotations() == ann and h2.getMetadata()["global"] == h.getMetadata()["global"]


def make_wav(path, channels, sampwidth, frames, rate=48000, pcm=None):
    rng = np.random.default_rng(5)
    if pcm is not None:
        pass
    elif sampwidth == 2:
        pcm = rng.integers(-30000, 30000, frames * channels, dtype=np.int16).astype("<i2").tobytes()
    else:
        pcm = rng.integers(0, 255, frames * channels, dtype=np.uint8).tobytes()
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels); w.setsampwidth(sampwidth); w.setframerate(rate); w.writeframes(pcm)
    return pcm


def make_float_wav(path, frames, extra_chunk=b""):
    x = np.random.default_rng(6).normal(size=2 * frames).astype("<f4").tobytes()
    fmt = struct.pack("<HHIIHH", 3, 2, 96000, 96000 * 8, 8, 32)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + extra_chunk + b"data" + struct.pack("<I", len(x)) + x
    path.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    return x


def test_wav_import_rules(tmp_path):
    # NonconformingDatasetHelper.fromWavFile :109-161
    pcm = make_wav(tmp_path / "s16.wav", 2, 2, 1000)
    h = sigmf.NonconformingDatasetHelper.fromWavFile(tmp_path / "s16.wav", 433e6)
    g, c = h.meta["global"], h.meta["captures"][0]
    assert (g["core:datatype"], c["core:header_bytes"], g["core:sample_rate"], c["core:frequency"]) == ("ci16_le", 44, 48000.0, 433e6)
    assert g["core:dataset"] == "s16.wav" and h.metaFile.endswith("s16.sigmf-meta")
    h.writeSigMfFile()
    buf = sigmf.SigMfHelper().load(h.metaFile).getDataBuffer()
    assert bytes(buf) == pcm
    make_wav(tmp_path / "u8.wav", 2, 1, 1000)
    assert sigmf.NonconformingDatasetHelper.fromWavFile(tmp_path / "u8.wav").meta["global"]["core:datatype"] == "cu8"
    make_wav(tmp_path / "mono.wav", 1, 2, 1000)
    assert sigmf.NonconformingDatasetHelper.fromWavFile(tmp_path / "mono.wav").meta["global"]["core:datatype"] == "ri16_le"
    x = make_float_wav(tmp_path / "f32.wav", 500, extra_chunk=b"LIST" + struct.pack("<I", 6) + b"abcdef")
    h = sigmf.NonconformingDatasetHelper.fromWavFile(tmp_path / "f32.wav")
    assert h.meta["global"]["core:datatype"] == "cf32_le"
    # header = file length - frames*frame size, rounded DOWN to a frame boundary (:130-133): 58 -> 56
    assert h.meta["captures"][0]["core:header_bytes"] == (os.path.getsize(tmp_path / "f32.wav") - len(x)) // 8 * 8
    with pytest.raises(ValueError):
        sigmf.NonconformingDatasetHelper.fromWavFile(tmp_path / "nope.wav")
    (tmp_path / "junk.wav").write_bytes(b"not a wav file at all")
    with pytest.raises(ValueError):
        sigmf.NonconformingDatasetHelper.fromWavFile(tmp_path / "junk.wav")


@pytest.mark.parametrize("name,exp", [("a.cs16", "ci16_le"), ("A.CI16", "ci16_le"), ("x.cf32", "cf32_le"), ("x.cf64", "cf64_le"),
                                      ("x.ci8", "ci8"), ("x.cu8", "cu8"), ("x.bin", "cf32_le")])
def test_guess_datatype_from_extension(name, exp):
    assert sigmf.NonconformingDatasetHelper.guessDatatypeFromExtension(name) == exp


def test_analyze_selection_params():
    # MainController.java:702-730
    s, w, f, d, fs2 = sigmf.analyze_selection_params(2.0e6, 100e6, 10_000_000, 100.2e6, 100.3e6, 500_000, 100_000)
    assert (s, w, d) == (490_000, 120_000, 16) and f == pytest.approx(0.125) and fs2 == 125_000.0
    s, w, f, d, _ = sigmf.analyze_selection_params(2.0e6, 100e6, 1_000_000, 99e6, 101.5e6, 5_000, 999_000)
    assert (s, w, d) == (0, 1_000_000, 1)                 # clamped at both ends; floor(fs/bw) == 0 -> 1
    assert sigmf.annotation_row_params(2.0e6, 100e6, 0.25, 0.01, 100.25e6, 125e3) == (500_000, 20_000, 0.125, 16, False)


@pytest.mark.gpu
def test_wav_mapping_through_the_engine(tmp_path, engine):
    """16-bit stereo WAV -> ci16_le with a 44-byte header: the mapped samples start 4-byte aligned only."""
    # three tones over noise (synth.recording), so that the frame has signal bins and a floor
    make_wav(tmp_path / "rec.wav", 2, 2, 1024 * 9 + 17, pcm=synth.recording(1024 * 9 + 17, "ci16_le", seed=31).tobytes())
    h = sigmf.NonconformingDatasetHelper.fromWavFile(tmp_path / "rec.wav")
    h.writeSigMfFile()
    sm = sigmf.SigMfHelper().load(h.metaFile)
    buf = sm.getDataBuffer()
    assert sm.total_samples == 1024 * 9 + 17
    got = engine.spectrogram(buf, sm.datatype, 1024, 10)              # reference framing; last frame past EOF
    ref = co.spectrogram(np.asarray(buf), sm.datatype, 0, 1024, 1024, "rect", 10)
    assert (got[9] == -150.0).all()
    check_db_parity(got[:9], ref[:9])


def test_multi_capture_segments(tmp_path):
    """Three captures, each preceded by its own header (non-conforming dataset): byte offsets, lengths and the
    mapped bytes of each capture; captures[0] is what the reference's getDataBuffer starts at."""
    dt, bps = "ci16_le", 4
    parts = [synth.recording(n, dt, seed=20 + i).tobytes() for i, n in enumerate((1000, 300, 777))]
    hdrs = [b"H" * 12, b"I" * 8, b"J" * 20]
    blob = b"".join(h + p for h, p in zip(hdrs, parts))
    starts = [0, 1000, 1300]
    meta = {"global": {"core:datatype": dt, "core:sample_rate": 1e6, "core:version": "1.0.0", "core:dataset": "x.bin"},
            "captures": [{"core:sample_start": s, "core:frequency": 1e6 * (i + 1), "core:header_bytes": len(h)}
                         for i, (s, h) in enumerate(zip(starts, hdrs))], "annotations": []}
    (tmp_path / "x.bin").write_bytes(blob)
    mp = tmp_path / "x.sigmf-meta"
    mp.write_text(json.dumps(meta))
    h = sigmf.SigMfHelper().load(mp)
    segs = h.capture_segments()
    assert [s[1] for s in segs] == starts and [s[2] for s in segs] == [1000, 300, 777]
    assert [s[3] for s in segs] == [1e6, 2e6, 3e6]
    off = 0
    for i, (hd, p) in enumerate(zip(hdrs, parts)):
        off += len(hd)
        assert segs[i][0] == off
        assert bytes(h.capture_buffer(i)) == p
        off += len(p)
    assert bytes(h.getDataBuffer()[: len(parts[0])]) == parts[0]          # reference view: past captures[0]'s header
    # single-capture conforming file: one segment covering the payload
    mp2 = write_sigmf(tmp_path, "y", dt, parts[0])
    h2 = sigmf.SigMfHelper().load(mp2)
    assert h2.capture_segments() == [(0, 0, 1000, 100e6)]
    # a truncated last capture is clipped to whole samples available
    (tmp_path / "x.bin").write_bytes(blob[:-5])
    assert sigmf.SigMfHelper().load(mp).capture_segments()[2][2] == 775
