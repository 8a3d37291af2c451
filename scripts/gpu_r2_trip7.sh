#!/bin/bash
# Round 2, GPU trip 7: sorted-sample cache, NVTX, weighted balance: parity + full-size parity + bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'])
print([ (c['ms_count'],c['ms_mesh']) for c in d['calls']])
PY
tail -3 gpurun_out/bench.err
echo "== tpcf step per rank"; timeout 600 python scripts/gpu_shardsim_stat.py 1,8 2> gpurun_out/streams.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d.items(): print(w, 'max', round(r['max_ms'],2), 'eff', round(r['predicted_efficiency'],3), [round(x,2) for x in r['per_rank_ms']])
"
