#!/bin/bash
# ncu full capture of the mean_delta_sigma kernel (config 5 shape, reduced galaxy count)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
C5_NGAL=${C5_NGAL:-200000} C5_NPTCL=${C5_NPTCL:-100000000} timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_dsq \
    python scripts/gpu_configs.py 5 > gpurun_out/prof_dsq.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_dsq.ncu-rep
