#!/bin/bash
# Fast3 occupancy / register variants on the RR count of the bench step and on config 4
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for L in "" variants/lib_f3_12_1.so variants/lib_f3_10_1.so; do
  echo "== lib ${L:-default (8 warps x 2 blocks, 128 registers)}"
  HTB_LIB_PATH=${L:+$PWD/$L} timeout 600 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
rb = synthetic.config_rbins()
for _ in range(3):
    hb.npairs_3d(ran, ran, rb, period=250.0)
print("RR ms_count", _lib.last_stats["ms_count"])
s = torch.from_numpy(np.random.RandomState(43).uniform(0, 1000.0, (10000000, 3))).cuda()
for _ in range(3):
    hb.npairs_3d(s, s, rb, period=1000.0)
print("config4 npairs ms_count", _lib.last_stats["ms_count"])
PY
done
