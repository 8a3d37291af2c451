#!/bin/bash
# Fine particle mesh of the cell-resolved delta-sigma kernel (config 5): HTB_M2 = mx,my (heuristic: 16,16).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for M in ${MS:-"" 12,12 20,20 24,24 32,32 16,32 32,16}; do
echo -n "HTB_M2=${M:-heuristic} : "
env ${M:+HTB_M2=$M} timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2> gpurun_out/bench_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5 step',round(d['ms_per_step'],2),d['breakdown_ms'],'evaluated',d['config']['pairs_evaluated_per_step'], d['delta_sigma'][:2])
" || tail -2 gpurun_out/bench_c5.err
done
