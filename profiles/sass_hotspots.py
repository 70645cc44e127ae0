"""Aggregate an `ncu -i X.ncu-rep --page source --csv --kernel-name K` export (SASS view): instruction and
stall-sample shares per opcode and per 500-instruction region of the kernel.
Usage: python profiles/sass_hotspots.py src.csv [region size] [section]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]      # one section per profiled launch
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if sec >= len(heads):
    print("no such section (kernel not in the report?)")
    sys.exit(0)
hi = heads[sec]
end = heads[sec + 1] if sec + 1 < len(heads) else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0].startswith("0x")]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot_s = sum(int(r[iS]) for r in data) or 1
tot_i = sum(int(r[iI]) for r in data) or 1
print("samples", tot_s, "warp instructions", tot_i, "SASS lines", len(data))
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iSrc])
    o = m.group(2).split(".")[0] if m else "?"
    op[o] += int(r[iI])
    ops[o] += int(r[iS])
for o, c in op.most_common(22):
    print(f"{o:10s} inst {c:>11d} {100 * c / tot_i:5.1f}%   samples {ops[o]:>7d} {100 * ops[o] / tot_s:5.1f}%")
print()
step = int(sys.argv[2]) if len(sys.argv) > 2 else 500
for k in range(0, len(data), step):
    ch = data[k:k + step]
    print(f"SASS {k:5d}+  inst {100 * sum(int(r[iI]) for r in ch) / tot_i:5.1f}%  samples "
          f"{100 * sum(int(r[iS]) for r in ch) / tot_s:5.1f}%   MMA {sum('MMA' in r[iSrc] for r in ch):3d}  "
          f"LDG {sum('LDG' in r[iSrc] for r in ch):3d}  LDS {sum('LDS' in r[iSrc] for r in ch):3d}  "
          f"BAR {sum('BAR' in r[iSrc] for r in ch):2d}  ATOM {sum('ATOM' in r[iSrc] or 'RED' in r[iSrc] for r in ch):2d}")
