#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "binq or xyz or smu or mxyz or perobj or wxy or marked or proj or per_object or projected or weighted or rp_pi" 2>&1 | tail -4
timeout 600 python scripts/gpu_generic.py 2>&1 | grep -v "fast)" | cut -c1-200
