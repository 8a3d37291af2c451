#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== per-rank counters"; timeout 600 python scripts/gpu_r2_rankpairs.py 2>&1 | tail -30
echo "== tpcf step per rank (three streams)"; timeout 600 python scripts/gpu_shardsim_stat.py 1,2,4,8 2> gpurun_out/streams.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d.items(): print(w, 'max', round(r['max_ms'],2), 'eff', round(r['predicted_efficiency'],3), [round(x,2) for x in r['per_rank_ms']])
"
echo "== c5 shards"; timeout 900 python scripts/gpu_shardsim.py c5 1,8 2> gpurun_out/shard_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d.items():
    if w.isdigit(): print(w, 'max', round(r['max_wall_ms'],2), 'eff', round(r.get('predicted_efficiency',0),3), [round(p['calls_total_count_mesh_ms'][0][1],1) for p in r['per_rank']], [round(p['calls_total_count_mesh_ms'][0][2],1) for p in r['per_rank']])
"
tail -3 gpurun_out/shard_c5.err
