#!/bin/bash
# Round 2, GPU trip 4: the rewritten MarkedQ + own-range fast path: parity suite, configs 1, 3, 4 timings.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== configs 1 3 4"; timeout 900 python scripts/gpu_configs.py 1 3 4 > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/configs.json'))
for k,v in d.items():
    for kk,vv in v.items():
        if isinstance(vv,dict) and 'ms_count' in vv:
            print(k,kk,'ms_count',vv['ms_count'],'ms_total',vv['ms_total'],'pairs',vv['pairs_evaluated'],'redone',vv['tiles_redone'],'tiles',vv['tiles'],'path',vv['path'])
    print(k,'ok',v.get('ok'))
PY
tail -3 gpurun_out/configs.err
