#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "device" 2>&1 | tail -5
timeout 600 python scripts/gpu_shardsim.py c5 1,8 > gpurun_out/shardsim_c5.json 2> gpurun_out/shardsim_c5.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/shardsim_c5.json"))
for w,v in d.items(): print("c5 world",w,"max_wall",v["max_wall_ms"],"rank0",v["per_rank"][0])
PY
timeout 600 python scripts/gpu_shardsim.py tpcf 1,8 > gpurun_out/shardsim_tpcf.json 2> gpurun_out/shardsim_tpcf.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/shardsim_tpcf.json"))
for w,v in d.items(): print("tpcf world",w,"max_wall",v["max_wall_ms"],"rank0",v["per_rank"][0])
PY
