#!/bin/bash
# One `ncu --set full` capture of k_integrate for a gpu_probe case, exported to small CSVs (gpurun merges at most 64 MiB
# back; a .ncu-rep with sources is ~25 MB): usage  tools/ncu_capture.sh <probe-case> <tag> [keep]
#   gpurun_out/<tag>_raw.csv      every metric of the launch (ncu --page raw --csv)
#   gpurun_out/<tag>_source.csv   per-SASS-instruction counters (ncu --page source --csv)
#   gpurun_out/<tag>.ncu-rep      only with `keep`
set -u
cd "${GRAFT_REPO_ROOT:-.}"
case_name=$1; tag=$2; keep=${3:-}
rep=/tmp/${tag}.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate -c 1 -f -o ${rep%.ncu-rep} \
    python tools/gpu_probe.py "$case_name" > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu $case_name rc=$?"
ncu -i $rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ls -la gpurun_out/${tag}_raw.csv gpurun_out/${tag}_source.csv
[ -n "$keep" ] && [ "$(du -sm $rep | cut -f1)" -lt 30 ] && cp $rep gpurun_out/
exit 0
