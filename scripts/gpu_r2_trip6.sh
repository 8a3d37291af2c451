#!/bin/bash
# Round 2, GPU trip 6: samples uploaded once + engine calls on separate streams: parity, bench (one stream vs three).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for ONE in 1 ""; do
echo "== bench HTB_ONE_STREAM=$ONE"; HTB_ONE_STREAM=$ONE timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/bench$ONE.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench$ONE.log').read().strip().splitlines()[-1])
print('step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'])
print([ (c['ms_count'],c['ms_mesh']) for c in d['calls']])
PY
tail -3 gpurun_out/bench.err
done
