#!/bin/bash
# Round 2, GPU trip 8: own-range chunks straight to the exact path, MarkedQ exact scan on the FP64 pipe, new K1: parity + configs.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_r2_isolate.sh
echo "== pytest gpu (parity)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
