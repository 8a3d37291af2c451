#!/bin/bash
# Round 2 ncu evidence: launch list of the bench command + one full-set capture of the dominant kernel of configs 2, 3, 4, 5.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-c5 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== Fast3 (bench RR launch)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count -s 2 -c 1 -f -o gpurun_out/prof_fast3 \
    python bench.py --steps 1 --warmup 0 --no-c5 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1; echo "rc=$?"
echo "== MarkedQ (config 4)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_markedq python scripts/gpu_configs.py 4 > gpurun_out/prof_markedq.log 2>&1; echo "rc=$?"
echo "== FastXYZ (config 3)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_fastxyz python scripts/gpu_configs.py 3 > gpurun_out/prof_fastxyz.log 2>&1; echo "rc=$?"
echo "== DSigmaR (config 5)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_dsr python scripts/gpu_configs.py 5 > gpurun_out/prof_dsr.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
