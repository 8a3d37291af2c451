#!/bin/bash
# Round 2, GPU trip 2: full-size reference parity (fixed ranges), bench (both arms), ncu captures of MarkedQ / FastXYZ.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (full size)"; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 900 --durations=0 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_full.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_ref.log
echo "== ncu MarkedQ"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_markedq python scripts/gpu_configs.py 4 > gpurun_out/prof_markedq.log 2>&1; echo "rc=$?"
echo "== ncu FastXYZ"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o gpurun_out/prof_fastxyz python scripts/gpu_configs.py 3 > gpurun_out/prof_fastxyz.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
