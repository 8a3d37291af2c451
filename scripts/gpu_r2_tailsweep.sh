#!/bin/bash
# A/B of the tail-phase slicing on one GPU: every rank's shard of the 8-GPU tpcf step run one after the other.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for T in ${TAILS:-0 2 4 8}; do
  echo "== HTB_TAIL_EIGHTHS=$T"
  HTB_TAIL_EIGHTHS=$T timeout 600 python scripts/gpu_shardsim.py tpcf 1,8 > gpurun_out/tail_$T.json 2> gpurun_out/tail_$T.err
  python - <<PY
import json
d=json.load(open('gpurun_out/tail_$T.json'))
for w in ('1','8'):
    r=d[w]
    print(w,'max_wall',round(r['max_wall_ms'],2),'RR count per rank',[round(p['calls_total_count_mesh_ms'][2][1],2) for p in r['per_rank']],'DD',[round(p['calls_total_count_mesh_ms'][0][1],2) for p in r['per_rank']],'DR',[round(p['calls_total_count_mesh_ms'][1][1],2) for p in r['per_rank']])
PY
done
echo "== pytest input step"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "input_step or device_estimator or asynchronous" 2>&1 | tail -5
