#!/bin/bash
# Wrapper for gpurun commands: runs "$@" from the repo root, then prunes gpurun_out/ below the 64 MiB that gpurun
# merges back (largest files go first) so that a stray .ncu-rep never costs the whole call's outputs.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bash -c "$*"
rc=$?
while [ "$(du -sm gpurun_out | cut -f1)" -ge 60 ]; do
    big=$(ls -S gpurun_out | head -n 1)
    echo "[gpu_job] pruning gpurun_out/$big ($(du -sm "gpurun_out/$big" | cut -f1) MiB)"
    rm -rf "gpurun_out/$big"
done
exit $rc
