#!/bin/bash
# Round 2, GPU trip 5: MarkedQ (8-byte entries), finer tail work items, DSigmaR renormalisation interval: parity, configs, bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== configs"; timeout 900 python scripts/gpu_configs.py ${CONFIGS:-1 3 4 5} > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/configs.json'))
for k,v in d.items():
    for kk,vv in v.items():
        if isinstance(vv,dict) and 'ms_count' in vv:
            print(k,kk,'ms_count',vv['ms_count'],'ms_mesh',vv['ms_mesh'],'ms_total',vv['ms_total'],'pairs',vv['pairs_evaluated'],'redone',vv['tiles_redone'],'tiles',vv['tiles'],'path',vv['path'])
    print(k,'ok',v.get('ok'), {a:b for a,b in v.items() if 'wall' in a})
PY
tail -3 gpurun_out/configs.err
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'])
print([ (c['ms_count'],c['ms_mesh']) for c in d['calls']])
PY
tail -3 gpurun_out/bench.err
