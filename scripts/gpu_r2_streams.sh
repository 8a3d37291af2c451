#!/bin/bash
# one stream vs three streams (+ early exit) on the shards of the 8-GPU tpcf step, all on one GPU
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for V in "HTB_ONE_STREAM=1" "HTB_ONE_STREAM=" "HTB_ONE_STREAM=1 HTB_EARLY_EXIT=1"; do
  echo "== $V"
  env $V timeout 600 python scripts/gpu_shardsim_stat.py 1,8 2> gpurun_out/streams.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d.items(): print(w, 'max', round(r['max_ms'],2), 'eff', round(r['predicted_efficiency'],3), [round(x,2) for x in r['per_rank_ms']])
"
  tail -2 gpurun_out/streams.err
done
