#!/bin/bash
# One-call capture of the wide-state kernels (csrc/engine_wide.cuh) on the GPU box, then the summaries here.
#   on the box (under gpurun):   bash profiles/capture_wide.sh box TAG [H]     -> gpurun_out/prof_wide_TAG.ncu-rep
#   in the container afterwards: bash profiles/capture_wide.sh read TAG        -> profiles/prof_wide_TAG_raw.csv,
#                                                                                 profiles/wide_TAG_{fwd,bwd}_{hotspots,lines}.txt
# The plain run comes first (a number printed under ncu is never a bench value; the capture only starts once the
# program has exited 0 without the profiler).
set -e
mode=$1; tag=$2; h=${3:-32}
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root"
if [ "$mode" = box ]; then
    mkdir -p gpurun_out
    timeout 60 python profiles/prof_step.py --h "$h" --layers 3 --steps 1
    timeout 150 ncu --set full --clock-control none --import-source on -k regex:wide -c 4 -f \
        -o "gpurun_out/prof_wide_$tag" python profiles/prof_step.py --h "$h" --layers 3 --steps 1 \
        > "gpurun_out/ncu_wide_$tag.log" 2>&1
    tail -2 "gpurun_out/ncu_wide_$tag.log"
else
    rep="gpurun_out/prof_wide_$tag.ncu-rep"
    ncu -i "$rep" --page raw --csv > "profiles/prof_wide_${tag}_raw.csv" 2>/dev/null
    for k in fwd bwd; do
        ncu -i "$rep" --page source --csv --kernel-name "${k}_wide_kernel" > "gpurun_out/src_${k}_$tag.csv" 2>/dev/null
        ncu -i "$rep" --page source --print-source cuda,sass --csv --kernel-name "${k}_wide_kernel" \
            > "gpurun_out/src_${k}_${tag}_cuda.csv" 2>/dev/null
        python profiles/sass_hotspots.py "gpurun_out/src_${k}_$tag.csv" 400 0 | awk '!/SASS/ || $4+0>0.9 || $6+0>0.9' \
            > "profiles/wide_${tag}_${k}_hotspots.txt"
        python profiles/cuda_line_hotspots.py "gpurun_out/src_${k}_${tag}_cuda.csv" 40 > "profiles/wide_${tag}_${k}_lines.txt"
    done
    python - "$tag" <<'PY'
import csv, sys
rows = list(csv.reader(open("profiles/prof_wide_%s_raw.csv" % sys.argv[1])))
hdr, data = rows[0], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"]
want += [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""),
              [r[i][:9] for r in data])
PY
fi
