"""Per-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`.
usage: python profiles/ncu_lines.py lines.csv <units> [top]   (units = frames / tiles per launch)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
c_inst, c_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
c_wf, c_wfi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
out, tot, tots = [], 0, 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[2] != "-":        # keep the per-line aggregate rows (Address == "-")
        continue
    try:
        n, s = int(r[c_inst]), int(r[c_samp])
        wf, wfi = int(r[c_wf]), int(r[c_wfi])
    except ValueError:
        continue
    tot += n
    tots += s
    out.append((n, s, wf, wfi, r[0], r[1].strip()[:100]))
out.sort(reverse=True)
print("warp instructions per unit: %.1f   samples: %d" % (tot / units, tots))
print("  inst/unit  samp%   smem-wf/unit (ideal)  line")
for n, s, wf, wfi, l, src in out[:top]:
    print("%9.1f %6.1f %9.1f %9.1f   L%-4s %s" % (n / units, 100.0 * s / max(tots, 1), wf / units, wfi / units, l, src))
