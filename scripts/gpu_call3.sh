#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("ms_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], [ (c["ms_count"], c["tiles"]) for c in d["calls"]], d["roofline"]["frac"])
PY
echo "== shardsim tpcf"; timeout 300 python scripts/gpu_shardsim.py tpcf > gpurun_out/shardsim_tpcf.json 2> gpurun_out/shardsim_tpcf.err; tail -5 gpurun_out/shardsim_tpcf.err
