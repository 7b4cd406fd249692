"""Hot straight-line segments of a kernel from an `ncu --page source --csv` dump: consecutive SASS instructions with
the same executed count are one segment (a loop body, a prologue ...); prints each segment's share of all executed
warp instructions and its opcode histogram.  This is how the staging batch of downconvert_kernel (42 % of the
instructions, 16 S2R + 33 ISETP per batch) and the library log10 / sqrt of the FP64 epilogue were found.

usage: ncu -i prof.ncu-rep --page source --csv -k regex:<kernel> > src.csv
       python profiles/ncu_segments.py src.csv [min_share_percent]
"""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().split("\n")
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
heads = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
end = heads[1] - 1 if len(heads) > 1 else len(lines)          # first table only (one launch)
rows = list(csv.reader(lines[:end]))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]


def opcode(r):
    t = r[ci["Source"]].split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


segs, start, prev = [], 0, None
for i, r in enumerate(body):
    n = int(r[ci["Instructions Executed"]])
    if n != prev:
        if prev is not None:
            segs.append((start, i - 1, prev))
        start, prev = i, n
segs.append((start, len(body) - 1, prev))
total = sum((b - a + 1) * n for a, b, n in segs)
print("%s\n%d SASS instructions, %d executed warp instructions" % (rows[0][1] if len(rows[0]) > 1 else "", len(body), total))
for a, b, n in segs:
    share = 100.0 * (b - a + 1) * n / max(total, 1)
    if share < min_share:
        continue
    ops = collections.Counter(opcode(r) for r in body[a:b + 1])
    print("  [%5d..%5d] x %-10d %4d instr  %5.1f %%  %s" % (a, b, n, b - a + 1, share,
          " ".join("%s:%d" % kv for kv in ops.most_common(8))))
