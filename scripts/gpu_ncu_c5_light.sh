#!/bin/bash
# lighter ncu capture of the mean_delta_sigma kernel (config 5): source counters + scheduler / warp-state / compute sections
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
C5_NGAL=${C5_NGAL:-1000000} timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis \
    --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_dsr \
    python scripts/gpu_configs.py 5 > gpurun_out/prof_dsr.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_dsr.ncu-rep
