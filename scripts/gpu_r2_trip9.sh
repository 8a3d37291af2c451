#!/bin/bash
# Round 2, GPU trip 9: fewer launches in the sort (one-block scan for small meshes, counters zeroed by the scan, pad by the
# scatter): small-call wall times, parity, c5 step.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== small calls"; timeout 600 python scripts/gpu_r2_small.py 2>&1 | tail -8
echo "== pytest gpu (parity)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== c5"; timeout 900 python bench.py --workload c5 --steps 3 --warmup 2 2> gpurun_out/bench_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5 step',d['ms_per_step'],'frac',d['roofline']['frac'],d['breakdown_ms'], d['delta_sigma'][:3])
"
