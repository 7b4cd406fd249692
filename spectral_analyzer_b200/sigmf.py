"""SigMF ingest -- the step that hands the engine its bytes.

Host-side mirror of (S/ = src/main/java/net/kcundercover/spectral_analyzer/ in the reference)

  SigMfHelper.load / getDataBuffer / getMetadata        S/sigmf/SigMfHelper.java:43-94, :123-136
  Global.getBytesPerSample                               S/sigmf/Global.java:67-79
  NonconformingDatasetHelper.fromWavFile / guessDatatypeFromExtension / writeSigMfFile
                                                         S/sigmf/NonconformingDatasetHelper.java:109-161, :196-210, :229-239
  annotation -> downconvert parameters                   S/controllers/MainController.java:696-730,
                                                         S/controllers/AnnotationController.java:329-335

with ONE deliberate difference: the data file is mapped whole with 64-bit offsets (numpy.memmap), not
clamped to Integer.MAX_VALUE bytes (SigMfHelper.java:78-82), so 4 GiB and 16 GiB recordings reach the
engine.  No arithmetic on samples happens here: the mapped bytes go to the C-ABI untouched.
"""
import json
import math
import os
import struct

import numpy as np


def bytes_per_sample(datatype):
    """Global.getBytesPerSample (Global.java:67-79), including its fallback of 8."""
    if datatype.startswith("cf32"):
        return 8
    if datatype.startswith("ci16"):
        return 4
    if datatype.startswith("cu8") or datatype.startswith("ci8"):
        return 2
    if datatype.startswith("cf64"):
        return 16
    return 8


class SigMfHelper:
    """S/sigmf/SigMfHelper.java.  load() parses the .sigmf-meta JSON and maps the data file read-only from
    core:header_bytes of the first capture; getDataBuffer() is that mapping (position 0 = sample 0)."""

    def __init__(self):
        self.metadata = None
        self.dataBuffer = None
        self.inputMeta = None
        self.dataPath = None

    def load(self, metaPath):
        metaPath = os.fspath(metaPath)
        with open(metaPath, "r", encoding="utf-8") as f:
            self.metadata = json.load(f)                               # :45
        g = self.metadata.get("global") or {}
        parent = os.path.dirname(metaPath)
        if g.get("core:dataset") is not None and parent != "":        # :49-52 non-conforming dataset
            dataPath = os.path.join(parent, g["core:dataset"])
        else:                                                          # :54-56
            dataPath = metaPath.replace(".sigmf-meta", ".sigmf-data")
        headerBytes = 0                                                # :59-67
        caps = self.metadata.get("captures") or []
        if caps and caps[0].get("core:header_bytes") is not None:
            headerBytes = int(caps[0]["core:header_bytes"])
        size = os.path.getsize(dataPath)                               # FileNotFoundException -> OSError
        available = max(0, size - headerBytes)                         # :76
        if "core:datatype" not in g:
            raise ValueError("global core:datatype missing")           # NullPointerException at :87 in the reference
        # the whole payload, 64-bit offsets (the reference clamps to 2 GiB - 1 here, :78-84)
        self.dataBuffer = (np.memmap(dataPath, dtype=np.uint8, mode="r", offset=headerBytes, shape=(available,))
                           if available else np.zeros(0, np.uint8))
        self.inputMeta = metaPath
        self.dataPath = dataPath
        return self

    # -- getters of the Java class
    def getMetadata(self):
        return self.metadata

    def getDataBuffer(self):
        return self.dataBuffer

    def getParsedAnnotations(self):
        return self.metadata.get("annotations") or []

    def getCurrentMetaFile(self):                                      # :103-115
        if self.inputMeta is None:
            return None
        if self.inputMeta.endswith(".sigmf-data"):
            return self.inputMeta.replace(".sigmf-data", ".sigmf-meta")
        return self.inputMeta

    def saveSigMF(self, annotationList):                               # :150-167
        self.metadata = {"global": self.metadata.get("global"), "captures": self.metadata.get("captures"),
                         "annotations": list(annotationList)}
        with open(self.getCurrentMetaFile(), "w", encoding="utf-8") as f:
            json.dump(self.metadata, f, indent=2)

    # -- what the engine needs
    @property
    def datatype(self):
        return self.metadata["global"]["core:datatype"]

    @property
    def sample_rate(self):
        return float(self.metadata["global"].get("core:sample_rate") or 1.0)

    @property
    def center_frequency(self):
        caps = self.metadata.get("captures") or []
        return float(caps[0].get("core:frequency") or 0.0) if caps else 0.0

    def capture_segments(self):
        """Beyond the reference (which honours captures[0] only, SigMfHelper.java:59-67): every capture of a
        multi-capture recording as (byte_offset, sample_start, n_samples, frequency), byte_offset counted from the
        start of the data FILE.  core:header_bytes of capture i is the number of non-sample bytes that precede its
        samples, so its samples start at sum(header_bytes[0..i]) + sample_start * bytes_per_sample; it ends where
        the next capture's header starts (or at the end of the file)."""
        caps = self.metadata.get("captures") or [{}]
        bps = bytes_per_sample(self.datatype)
        size = os.path.getsize(self.dataPath)
        segs, hdr = [], 0
        for i, c in enumerate(caps):
            hdr += int(c.get("core:header_bytes") or 0)
            s0 = int(c.get("core:sample_start") or 0)
            off = hdr + s0 * bps
            if i + 1 < len(caps):
                n = int(caps[i + 1].get("core:sample_start") or 0) - s0
            else:
                n = max(0, size - off) // bps
            n = max(0, min(n, max(0, size - off) // bps))
            segs.append((off, s0, n, float(c.get("core:frequency") or 0.0)))
        return segs

    def capture_buffer(self, index):
        """The bytes of one capture, mapped read-only with 64-bit offsets (position 0 = its first sample): what
        the engine is handed for that capture."""
        off, _, n, _ = self.capture_segments()[index]
        nbytes = n * bytes_per_sample(self.datatype)
        if nbytes == 0:
            return np.zeros(0, np.uint8)
        return np.memmap(self.dataPath, dtype=np.uint8, mode="r", offset=off, shape=(nbytes,))

    @property
    def total_samples(self):                                           # MainController.java:603-605
        return len(self.dataBuffer) // bytes_per_sample(self.datatype)


def _parse_wav(path):
    """RIFF/WAVE chunk walk: returns (format_tag, channels, sample_rate, bits, frame_bytes, data_bytes)."""
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError("not a RIFF/WAVE file")                   # UnsupportedAudioFileException
        fmt = None
        while True:
            ck = f.read(8)
            if len(ck) < 8:
                raise ValueError("WAV file has no data chunk")
            cid, csz = ck[:4], struct.unpack("<I", ck[4:])[0]
            if cid == b"fmt ":
                body = f.read(csz + (csz & 1))
                tag, ch, rate, _brate, align, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == 0xFFFE and csz >= 26:                        # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, ch, rate, bits, align)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError("WAV data chunk before fmt chunk")
                remaining = os.path.getsize(path) - f.tell()
                return fmt + (min(csz, remaining),)
            else:
                f.seek(csz + (csz & 1), 1)


class NonconformingDatasetHelper:
    """S/sigmf/NonconformingDatasetHelper.java: SigMF metadata for raw / WAV files (pure metadata)."""

    def __init__(self, dataFile, sampleRate, centerFreq, datatype, headerBytes):
        dataFile = os.fspath(dataFile)
        base = os.path.basename(dataFile)
        dot = base.rfind(".")
        stem = base[:dot] if dot > 0 else base                         # buildMetaFile :219-226
        self.metaFile = os.path.join(os.path.dirname(dataFile), stem + ".sigmf-meta")
        self.meta = {
            "global": {"core:datatype": datatype, "core:sample_rate": float(sampleRate), "core:version": "1.0.0",
                       "core:dataset": base},
            "captures": [{"core:sample_start": 0, "core:frequency": float(centerFreq),
                          "core:header_bytes": int(headerBytes)}],
            "annotations": [],
        }

    @staticmethod
    def fromWavFile(wavFile, defaultCenterFreq=0):
        """:109-161.  channels 2 -> 'c', 1 -> 'r'; float32 -> f32_le, 16 bit -> i16_le, 8 bit -> u8 (no byte
        order), anything else falls back to f32_le; header_bytes = file length - frames * frame size,
        rounded down to a frame boundary (:130-133)."""
        wavFile = os.fspath(wavFile)
        if not os.path.exists(wavFile):
            raise ValueError("Target WAV file reference must exist on disk.")
        tag, channels, rate, bits, frame, data_bytes = _parse_wav(wavFile)
        if channels > 2:
            raise ValueError("Unsupported WAV format: more than 2 channels is not supported for SDR data. "
                             "Found %d channels." % channels)
        prefix = "c" if channels == 2 else "r"
        frames = data_bytes // frame if frame else 0
        headerBytes = os.path.getsize(wavFile) - frames * frame
        if frame and headerBytes % frame != 0:
            headerBytes -= headerBytes % frame
        if tag == 3 and bits == 32:
            dtype = prefix + "f32_le"
        elif bits == 16:
            dtype = prefix + "i16_le"
        elif bits == 8:
            dtype = prefix + "u8"
        else:
            dtype = prefix + "f32_le"
        return NonconformingDatasetHelper(wavFile, rate, defaultCenterFreq, dtype, headerBytes)

    @staticmethod
    def guessDatatypeFromExtension(filename):                          # :196-210
        lower = filename.lower()
        if lower.endswith(".cs16") or lower.endswith(".ci16"):
            return "ci16_le"
        if lower.endswith(".cf32"):
            return "cf32_le"
        if lower.endswith(".cf64"):
            return "cf64_le"
        if lower.endswith(".ci8"):
            return "ci8"
        if lower.endswith(".cu8"):
            return "cu8"
        return "cf32_le"

    def getMetaFilePath(self):
        return os.path.abspath(self.metaFile)

    def writeSigMfFile(self):                                          # :229-239
        with open(self.metaFile, "w", encoding="utf-8") as f:
            json.dump(self.meta, f, indent=2)
        return self.metaFile


def analyze_selection_params(inputFs, inputFc, totalSamples, selectionFreqLow, selectionFreqHigh,
                             selectionStartSample, selectionWidthSamples):
    """MainController.handleAnalyzeSelection :702-730: (targetStart, targetWidth, freqOff, down, targetFs)."""
    currBw = (selectionFreqHigh - selectionFreqLow) * 1.2
    center = (selectionFreqHigh + selectionFreqLow) / 2.0 - inputFc
    ext = int(selectionWidthSamples * 0.1)
    targetStart = 0 if selectionStartSample - ext < 0 else selectionStartSample - ext
    if selectionStartSample + selectionWidthSamples * 1.1 > totalSamples:
        targetWidth = totalSamples - targetStart
    else:
        targetWidth = int(selectionWidthSamples * 1.1) + (selectionStartSample - targetStart)
    tmpDown = int(math.floor(inputFs / currBw))
    down = 1 if tmpDown == 0 else tmpDown
    return targetStart, targetWidth, center / inputFs, down, inputFs / down


def annotation_row_params(sampleRate, inputFc, startTime, duration, centerFreq, bandwidth):
    """AnnotationController.executeCapability :329-337 (fast = false): the tuple Engine.downconvert_psd_batch
    takes, (start_sample, count, freq_off, down, fast)."""
    down = int(math.floor(sampleRate / bandwidth))
    return int(startTime * sampleRate), int(duration * sampleRate), (centerFreq - inputFc) / sampleRate, down, False
