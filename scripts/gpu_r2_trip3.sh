#!/bin/bash
# Round 2, GPU trip 3: asynchronous engine calls + K3 on the device (parity suite, bench).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
