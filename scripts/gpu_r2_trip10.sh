#!/bin/bash
# Round 2, GPU trip 10: tile list built by one warp per column: parity, bench, configs, small calls.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (parity)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for T in 1 ""; do
echo "== HTB_TILES_BY_THREAD=${T:-unset}"
env ${T:+HTB_TILES_BY_THREAD=$T} FZS=32 bash scripts/gpu_r2_refine.sh | grep -v "^=="
env ${T:+HTB_TILES_BY_THREAD=$T} timeout 900 python bench.py --workload c5 --steps 3 --warmup 2 2> gpurun_out/bench_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5 step',d['ms_per_step'],d['breakdown_ms'])
"
done
echo "== small calls"; timeout 600 python scripts/gpu_r2_small.py 2>&1 | tail -6
