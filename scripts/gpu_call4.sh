#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (delta sigma)"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "delta_sigma or ds_ or shards or slices" > gpurun_out/pytest_ds.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_ds.log
echo "== c5"; timeout 600 python scripts/gpu_shardsim.py c5 1 > gpurun_out/shardsim_c5.json 2> gpurun_out/shardsim_c5.err; tail -3 gpurun_out/shardsim_c5.err; cat gpurun_out/shardsim_c5.json | cut -c1-400
HTB_NO_DSR=1 timeout 600 python scripts/gpu_shardsim.py c5 1 2>&1 | tail -2 | cut -c1-300
