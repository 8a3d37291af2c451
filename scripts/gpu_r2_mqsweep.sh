#!/bin/bash
# MarkedQ occupancy / queue-depth variants on config 4
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for L in "" variants/lib_mq_4_3_32.so variants/lib_mq_8_2_32.so variants/lib_mq_4_4_32.so; do
  echo "== lib ${L:-default (4 warps x 3 blocks, 48 entries)}"
  HTB_LIB_PATH=${L:+$PWD/$L} timeout 600 python scripts/gpu_configs.py 4 2> gpurun_out/mq.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())['config4']
print('marked ms_count', d['marked_stats']['ms_count'], 'npairs ms_count', d['npairs_stats']['ms_count'], 'ok', d['ok'])
"
  tail -2 gpurun_out/mq.err
done
