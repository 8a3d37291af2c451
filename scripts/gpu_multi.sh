#!/bin/bash
# multi-GPU bench (torchrun, one rank per GPU): the driver's workload and the config-5 scaling workload
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== tpcf N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -2 gpurun_out/bench_n$N.err; tail -1 gpurun_out/bench_n$N.log | cut -c1-250
[ -n "${SKIP_C5:-}" ] && exit 0
echo "== c5 N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$N.log 2> gpurun_out/bench_c5_n$N.err; echo "rc=$?"; tail -2 gpurun_out/bench_c5_n$N.err; tail -1 gpurun_out/bench_c5_n$N.log | cut -c1-250
