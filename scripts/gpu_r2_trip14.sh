#!/bin/bash
# Round 2, GPU trip 14: per-sample upload events in DeviceStatistic (DD runs while the randoms are being copied): device-statistics
# tests, bench (e2e arms).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_statistic or device_estimator or tpcf or wp or rp_pi or asynchronous" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-cpu-baseline 2> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tpcf step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'frac',d['roofline']['frac'])
"
tail -2 gpurun_out/bench.err
