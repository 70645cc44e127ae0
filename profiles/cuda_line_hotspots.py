"""Per-CUDA-line totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name K`:
stall samples and executed warp instructions of the first profiled launch, top lines per file.
Usage: python profiles/cuda_line_hotspots.py src_cuda.csv [top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
seen, out, fname, hdr = set(), {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1]
        if fname in seen:           # second launch starts: stop
            break
        seen.add(fname)
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    try:
        out[(fname, int(r[0]))] = (int(r[iS]), int(r[iI]), r[1].strip()[:110])
    except ValueError:
        pass
ts = sum(v[0] for v in out.values()) or 1
ti = sum(v[1] for v in out.values()) or 1
print("samples", ts, "warp instructions", ti)
for (f, ln), (s, i, src) in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * s / ts:5.1f}% smp {100 * i / ti:5.1f}% inst  {(f or "?").split("/")[-1]}:{ln}  {src}")
