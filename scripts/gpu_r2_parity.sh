#!/bin/bash
# Round 2, first GPU trip: the parity suite with the tightened delta-sigma gates (achieved errors recorded),
# then the full-size reference-parity tests.  Outputs under gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/ds_errors.jsonl
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== pytest gpu (parity)"; HTB_RECORD_ERRORS=$PWD/gpurun_out/ds_errors.jsonl timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== pytest gpu (full size)"; HTB_RECORD_ERRORS=$PWD/gpurun_out/ds_errors.jsonl timeout 2400 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 1500 --durations=0 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_full.log
