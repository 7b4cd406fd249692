"""Per-frame dynamic instruction mix and stall summary from an `ncu --page source --csv` dump.
usage: python profiles/ncu_mix.py src.csv <frames>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2])
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ops, stalls, tot, samples = collections.Counter(), collections.Counter(), 0, 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    n = int(r[ci["Instructions Executed"]])
    t = r[ci["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += n
    tot += n
    for s in stall_cols:
        stalls[s] += int(r[ci[s]])
    samples += int(r[ci["# Samples"]])
print("warp instructions per frame: %.1f" % (tot / frames))
for k, v in ops.most_common(28):
    print("  %-10s %8.1f" % (k, v / frames))
print("stall samples (all): %d" % samples)
for k, v in stalls.most_common(10):
    print("  %-26s %5.1f%%" % (k, 100.0 * v / max(samples, 1)))
