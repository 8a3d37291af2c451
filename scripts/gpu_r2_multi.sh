#!/bin/bash
# multi-GPU validation + bench at N ranks (gpurun --gpus N)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
echo "== multi-GPU check N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_r2_multi.py 2>&1 | grep -v "^W\|OMP_NUM" | tail -15
echo "== bench N=$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print('N=$N step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'value',d['value'])
    c=d.get('c5')
    if c: print('c5 step',c['ms_per_step'],'e2e',c['e2e']['ms_per_step'],'frac',c['roofline']['frac'])
except Exception as e:
    print('no bench line', e)
PY
tail -5 gpurun_out/bench_n$N.err
