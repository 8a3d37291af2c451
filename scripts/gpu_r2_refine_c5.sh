#!/bin/bash
# Fine meshes of the cell-resolved delta-sigma kernel (config 5): HTB_M2 = mx,my of the particle mesh (heuristic 16,16),
# HTB_M1 = mx,my of the galaxy mesh (heuristic 4,15).   usage: VAR=HTB_M1 MS="..." gpu_r2_refine_c5.sh
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
VAR=${VAR:-HTB_M1}
for M in ${MS:-heuristic 3,15 6,15 8,15 4,8 4,24 6,24}; do
echo -n "$VAR=$M : "
if [ "$M" = "heuristic" ]; then E=""; else E="$VAR=$M"; fi
env $E timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2> gpurun_out/bench_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5 step',round(d['ms_per_step'],2),d['breakdown_ms'],'evaluated',d['config']['pairs_evaluated_per_step'], d['delta_sigma'][:2])
" || tail -2 gpurun_out/bench_c5.err
done
