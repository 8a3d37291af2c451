#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for V in "" "HTB_ONE_PASS_SORT=1"; do
echo "== c5 bench $V"; env $V timeout 900 python bench.py --workload c5 --steps 3 --warmup 2 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c5.log').read().strip().splitlines()[-1])
print('c5 step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],d['breakdown_ms'])
PY
tail -3 gpurun_out/bench_c5.err
done
