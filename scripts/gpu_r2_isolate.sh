#!/bin/bash
# Which change moves the config-4 / config-3 kernels?  Same script, one library per line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for L in ${LIBS:-default}; do
[ "$L" = "default" ] && L=""
echo "== lib ${L:-default}"
HTB_LIB_PATH=${L:+$PWD/$L} timeout 600 python scripts/gpu_configs.py 4 2> gpurun_out/cfg4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['config4']
m,n=d['marked_stats'],d['npairs_stats']
print('config4 ok',d['ok'],'marked',round(m['ms_count'],2),'mesh',round(m['ms_mesh'],2),'evaluated',m['pairs_evaluated'],'redone',m['tiles_redone'],'| npairs',round(n['ms_count'],2),'mesh',round(n['ms_mesh'],2),'evaluated',n['pairs_evaluated'],'tiles',n['tiles'])
"
HTB_LIB_PATH=${L:+$PWD/$L} timeout 600 python scripts/gpu_configs.py 3 2> gpurun_out/cfg3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['config3']
m=d['stats']
print('config3 ok',d['ok'],'xy_z',round(m['ms_count'],2),'mesh',round(m['ms_mesh'],2),'evaluated',m['pairs_evaluated'],'redone',m['tiles_redone'])
"
HTB_LIB_PATH=${L:+$PWD/$L} timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-cpu-baseline 2> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tpcf step',d['ms_per_step'],'frac',d['roofline']['frac'],[ (round(c['ms_count'],2),round(c['ms_mesh'],2)) for c in d['calls']])
"
done
