#!/bin/bash
# Round 2, GPU trip 13: early asynchronous upload in mean_delta_sigma (one GPU, large host samples): delta-sigma parity cases,
# the full-size config-5 comparison, c5 bench (e2e).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ds_ or delta_sigma or sigma" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "config5" 2>&1 | tail -2
timeout 900 python bench.py --workload c5 --steps 3 --warmup 2 2> gpurun_out/bench_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5 step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['breakdown_ms'])
"
