#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5shard_launches.csv python scripts/gpu_ncu_c5shard.py > gpurun_out/c5shard.log 2>&1
echo rc=$?; tail -2 gpurun_out/c5shard.log
