#!/bin/bash
# Round 2, GPU trip 12: single-block set-up kernels at 256 threads (they must fit beside a resident persistent kernel):
# 8 ranks replayed on one GPU, timeline, parity.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python scripts/gpu_shardsim_stat.py 1,2,4,8 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d.items(): print(w, 'max', round(r['max_ms'],2), 'eff', round(r['predicted_efficiency'],3), [round(x,2) for x in r['per_rank_ms']])
"
timeout 600 python scripts/gpu_r2_timeline.py 2>&1 | tail -14
echo "== pytest gpu (parity)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
