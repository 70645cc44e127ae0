"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) count, average,
maximum and share of the summed device time.  Usage: python profiles/summarize_launches.py launches.csv [skip]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, gi, vi, ui = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
stat = collections.OrderedDict()
total = 0.0
n = 0
for r in rows[start + 1 + skip:]:
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    name = re.sub(r"^eng::", "", name)
    key = "%s grid=%s" % (name[:60], r[gi])
    s = stat.setdefault(key, [0, 0.0, 0.0])
    s[0] += 1
    s[1] += us
    s[2] = max(s[2], us)
    total += us
    n += 1
print("launches %d, summed device time %.3f ms (cold caches, serialised -> compare SHARES)" % (n, total / 1e3))
print("%-84s %4s %9s %9s %7s" % ("kernel", "n", "avg_us", "max_us", "share"))
for k, s in sorted(stat.items(), key=lambda kv: -kv[1][1]):
    print("%-84s %4d %9.1f %9.1f %6.1f%%" % (k, s[0], s[1] / s[0], s[2], 100 * s[1] / total))
