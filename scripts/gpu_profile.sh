#!/bin/bash
# ncu evidence for the dominant kernel: launch list of a short bench + one full-set capture of the RR launch.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${PROF_RANDOMS:-2000000}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --randoms $R --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_count -s 2 -c 1 -f -o gpurun_out/prof_fast3 \
    python bench.py --steps 1 --warmup 0 --randoms $R --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
