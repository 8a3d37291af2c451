#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== shardsim tpcf"; timeout 300 python scripts/gpu_shardsim.py tpcf 1,8 > gpurun_out/shardsim_tpcf.json 2> gpurun_out/shardsim_tpcf.err; tail -2 gpurun_out/shardsim_tpcf.err
echo "== shardsim c5"; timeout 600 python scripts/gpu_shardsim.py c5 1,2,4,8 > gpurun_out/shardsim_c5.json 2> gpurun_out/shardsim_c5.err; tail -4 gpurun_out/shardsim_c5.err
