#!/bin/bash
# One full-set ncu capture of the RR launch of the dominant kernel (no launch list).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${PROF_RANDOMS:-5000000}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_count -s 2 -c 1 -f -o gpurun_out/prof_fast3 \
    python bench.py --steps 1 --warmup 0 --randoms $R --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
