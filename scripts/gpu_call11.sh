#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q --timeout 600 -k "staged or per_object or perobj or delta_sigma or ds_" 2>&1 | tail -4
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5_n1.log 2> gpurun_out/bench_c5_n1.err; echo rc=$?; tail -1 gpurun_out/bench_c5_n1.log | cut -c1-400
timeout 300 python scripts/gpu_8f.py 2>&1 | grep "per_object" | cut -c1-200
