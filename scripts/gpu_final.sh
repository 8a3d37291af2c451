#!/bin/bash
# Round-end evidence on one B200: smoke, parity tests, bench (both arms), full-size configs, ncu launch list + full capture.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
echo "== configs"; timeout 900 python scripts/gpu_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; echo "rc=$?"; tail -2 gpurun_out/configs.err
echo "== generic"; timeout 600 python scripts/gpu_generic.py > gpurun_out/generic.log 2>&1; echo "rc=$?"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full capture (Fast3 RR launch)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_count -s 2 -c 1 -f -o gpurun_out/prof_fast3 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_fast3.ncu-rep
echo "== 8f counters"; timeout 600 python scripts/gpu_8f.py > gpurun_out/8f.log 2>&1; echo "rc=$?"
echo "== ncu full capture (BinQ rp_pi)"; bash scripts/gpu_ncu_binq.sh | tail -2
