#!/bin/bash
# Round 2, GPU trip 11: device-side kernel spans in the bench line; parity; async tests.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
r=d['roofline']
print('step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'frac',r['frac'],'kernel_ms',r['kernel_ms'],r['kernel'],r['rr_launch'])
PY
tail -3 gpurun_out/bench.err
echo "== pytest gpu (parity)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
