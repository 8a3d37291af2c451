#!/bin/bash
# compute-sanitizer memcheck of selected GPU tests.  usage: gpu_sanitize_one.sh <pytest -k expression> [tool]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TOOL=${2:-memcheck}
timeout 1200 compute-sanitizer --tool $TOOL --print-limit 5 --launch-timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$1" > gpurun_out/sanitize_$TOOL.log 2>&1
echo "rc=$?"; grep -v "^$" gpurun_out/sanitize_$TOOL.log | head -60
