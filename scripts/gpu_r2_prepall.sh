#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for V in "HTB_PREPARE_SMALL=1"; do
  echo "== $V"
  env $V timeout 600 python scripts/gpu_r2_timeline.py 2>&1 | tail -16
  env $V timeout 600 python scripts/gpu_shardsim_stat.py 1,8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d.items(): print(w, 'max', round(r['max_ms'],2), 'eff', round(r['predicted_efficiency'],3), [round(x,2) for x in r['per_rank_ms']])
"
done
