#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
SWEEP_M2="default" bash scripts/gpu_sweep.sh
mkdir -p v2; mv variants/lib_base.so v2/; rm variants/*.so; mv v2/lib_base.so variants/
SWEEP_M2="3,3,8 4,4,6 4,4,12 5,5,8 3,4,8 4,4,16" bash scripts/gpu_sweep.sh
