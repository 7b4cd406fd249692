#!/usr/bin/env python
"""Downconvert every annotation of a SigMF recording and take its Welch PSD, in one batched GPU call per PSD size.

    python tools/annotations_psd.py capture.sigmf-meta out.npz [--psd-nfft 8192] [--fast]

What the annotation table's "analyze" capability does row by row (AnnotationController.executeCapability,
S/controllers/AnnotationController.java:321-360: start/duration/centre/bandwidth -> extractAndDownConvert -> result)
plus the PSD of the Analysis dialog (AnalysisDialogController.java:303-313), for all annotations at once: only the
annotated spans cross PCIe.  out.npz holds, per annotation i: psd_db_i (dB/Hz, fft-shifted), freq_hz_i (centred on
the annotation's centre frequency), and the arrays start_sample / count / down / freq_off / psd_nfft / label.

The PSD size is --psd-nfft when the decimated annotation is at least that long, else the largest power of two that
fits (the reference passes the length itself to JDSP, :303-313; this engine transforms powers of two).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_analyzer_b200 import sigmf            # noqa: E402


def annotation_rows(helper, fast=False):
    """(start_sample, count, freq_off, down, fast) + (label, centre Hz) per SigMF annotation that names a band."""
    fs, fc, total = helper.sample_rate, helper.center_frequency, helper.total_samples
    rows, info = [], []
    for a in helper.getParsedAnnotations():
        lo, hi = a.get("core:freq_lower_edge"), a.get("core:freq_upper_edge")
        s0, cnt = int(a.get("core:sample_start") or 0), int(a.get("core:sample_count") or 0)
        if lo is None or hi is None or hi <= lo or cnt <= 0 or s0 >= total:
            continue
        cnt = min(cnt, total - s0)
        centre, bw = 0.5 * (lo + hi), hi - lo
        start, count, f_off, down, _ = sigmf.annotation_row_params(fs, fc, s0 / fs, cnt / fs, centre, bw)
        down = max(1, down)
        count = min(count, total - start)
        if count // down < 64:
            continue                                   # shorter than the smallest PSD this engine computes
        rows.append((start, count, f_off, down, bool(fast)))
        info.append((str(a.get("core:label") or a.get("core:description") or ""), centre))
    return rows, info


def psd_size(m, want):
    n = want
    while n > m:
        n //= 2
    return n


def run(engine, helper, psd_nfft=8192, fast=False):
    rows, info = annotation_rows(helper, fast)
    sizes = [psd_size(c // d, psd_nfft) for (_, c, _, d, _) in rows]
    out = {"start_sample": np.array([r[0] for r in rows], np.int64), "count": np.array([r[1] for r in rows], np.int64),
           "freq_off": np.array([r[2] for r in rows]), "down": np.array([r[3] for r in rows], np.int64),
           "psd_nfft": np.array(sizes, np.int64), "label": np.array([i[0] for i in info])}
    fs = helper.sample_rate
    for n in sorted(set(sizes)):
        idx = [i for i, s in enumerate(sizes) if s == n]
        _, psd = engine.downconvert_psd_batch(helper.getDataBuffer(), helper.datatype, fs, [rows[i] for i in idx],
                                              psd_nfft=n, want_iq=False)
        for j, i in enumerate(idx):
            fs_i = fs / rows[i][3]
            out["psd_db_%d" % i] = psd[j]
            out["freq_hz_%d" % i] = (np.arange(n) - n // 2) * fs_i / n + info[i][1]
    return out, rows, info


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("meta")
    ap.add_argument("npz")
    ap.add_argument("--psd-nfft", type=int, default=8192)
    ap.add_argument("--fast", action="store_true", help="moving-average polyphase form instead of the windowed-sinc low-pass")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    if args.psd_nfft < 64 or args.psd_nfft > 16384 or args.psd_nfft & (args.psd_nfft - 1):
        ap.error("--psd-nfft must be a power of two in 64..16384")
    import spectral_analyzer_b200 as sa
    h = sigmf.SigMfHelper().load(args.meta)
    eng = sa.Engine(args.device)                       # raises without a B200: there is no CPU fallback
    try:
        out, rows, info = run(eng, h, args.psd_nfft, args.fast)
    finally:
        eng.close()
    np.savez(args.npz, **out)
    for i, (r, inf) in enumerate(zip(rows, info)):
        p = out["psd_db_%d" % i]
        print("%3d %-16s start %12d count %10d down %4d  peak %8.2f dB/Hz at %.6g Hz" % (
            i, inf[0][:16], r[0], r[1], r[3], float(np.nanmax(p)), float(out["freq_hz_%d" % i][int(np.nanargmax(p))])))
    print("%d annotations -> %s" % (len(rows), args.npz))


if __name__ == "__main__":
    main()
