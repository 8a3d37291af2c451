#!/bin/bash
# Fine cells along the fast dimension: HTB_FZ = search length / cell height (round 1: 8).  Bench launches, configs 3 / 4, config 1.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for FZ in ${FZS:-8 16 24 32 48}; do
echo "== HTB_FZ=$FZ"
HTB_FZ=$FZ timeout 600 python bench.py --steps 3 --warmup 2 --no-c5 --no-cpu-baseline 2> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tpcf step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),[ (round(c['ms_count'],2),round(c['ms_mesh'],2),c['refine1'],c['refine2'],c['tiles_redone']) for c in d['calls']], 'evaluated', d['config']['pairs_evaluated_per_step'])
" || tail -2 gpurun_out/bench.err
HTB_FZ=$FZ timeout 600 python scripts/gpu_configs.py 1 3 4 2> gpurun_out/cfg.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['config1']; print('config1 ok',c['ok'],'wall',round(c['wall_s']*1e3,3),'count',round(c['stats']['ms_count'],3),'mesh',round(c['stats']['ms_mesh'],3),c['stats']['refine2'])
c=d['config3']; m=c['stats']; print('config3 ok',c['ok'],'xy_z',round(m['ms_count'],2),'mesh',round(m['ms_mesh'],2),m['refine2'],'evaluated',m['pairs_evaluated'])
c=d['config4']; m,n=c['marked_stats'],c['npairs_stats']; print('config4 ok',c['ok'],'marked',round(m['ms_count'],2),'mesh',round(m['ms_mesh'],2),m['refine2'],'| npairs',round(n['ms_count'],2),'mesh',round(n['ms_mesh'],2),'evaluated',n['pairs_evaluated'])
" || tail -2 gpurun_out/cfg.err
done
