#!/bin/bash
# block size of the Fast3 kernels (warps are independent: smaller blocks retire - and free their SM slot for the next
# kernel - earlier): 8 x 2, 4 x 4, 2 x 8 warps x blocks per SM, on the shards of the 8-GPU tpcf step and at N = 1
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for L in "" variants/lib_w4.so variants/lib_w2.so; do
  echo "== lib ${L:-default (8 warps x 2 blocks)}"
  HTB_LIB_PATH=${L:+$PWD/$L} timeout 600 python scripts/gpu_shardsim_stat.py 1,8 2> gpurun_out/streams.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for w,r in d.items(): print(w, 'max', round(r['max_ms'],2), 'eff', round(r['predicted_efficiency'],3), [round(x,2) for x in r['per_rank_ms']])
"
  tail -2 gpurun_out/streams.err
done
