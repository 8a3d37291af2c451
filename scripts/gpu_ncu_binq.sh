#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_binq \
    python scripts/gpu_binq_case.py xyz > gpurun_out/prof_binq.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_binq.log; ls -la gpurun_out/prof_binq.ncu-rep
