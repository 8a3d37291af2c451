#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { timeout 300 python scripts/gpu_shardsim.py c5 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())['1']['per_rank'][0]['calls_total_count_mesh_ms'][0]
print('count %.1f mesh %.1f' % (d[1], d[2]))"; }
for lib in variants/lib_*.so; do echo "== $lib"; HTB_LIB_PATH=$PWD/$lib run; done
for m2 in 8,8 12,12 16,8 16,16 24,24; do echo "== M2=$m2"; HTB_M2=$m2 run; done
for m1 in 4,12 5,16 8,16 8,24 10,32; do echo "== M1=$m1"; HTB_M1=$m1 run; done
