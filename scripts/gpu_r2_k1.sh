#!/bin/bash
# Round 2: K1 (cell assignment + counting sort) before / after the exact-floor digitize and the phase-batched scatter;
# parity suite on the new library; one ncu --set full capture of k_assign / k_scatter on the 1e8-particle sample.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for L in variants/lib_oldmesh.so ""; do
echo "== lib ${L:-default}"; HTB_LIB_PATH=${L:+$PWD/$L} timeout 900 python bench.py --workload c5 --steps 3 --warmup 2 2> gpurun_out/bench_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5 step',d['ms_per_step'],'frac',d['roofline']['frac'],d['breakdown_ms'], d['delta_sigma'][:3])
"
HTB_LIB_PATH=${L:+$PWD/$L} timeout 600 python scripts/gpu_configs.py 4 2> gpurun_out/cfg4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['config4']
print('config4 ok',d['ok'],'marked ms_count',d['marked_stats']['ms_count'],'mesh',d['marked_stats']['ms_mesh'],'npairs ms_count',d['npairs_stats']['ms_count'],'mesh',d['npairs_stats']['ms_mesh'])
"
done
echo "== pytest gpu (parity)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench tpcf"; timeout 900 python bench.py --steps 5 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['ms_per_step'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'])
print([ (c['ms_count'],c['ms_mesh']) for c in d['calls']])
PY
echo "== ncu K1"; C5_NGAL=20000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_scatter' -c 4 -f -o gpurun_out/prof_k1 \
    python scripts/gpu_configs.py 5 > gpurun_out/prof_k1.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_k1.ncu-rep
