"""One config-5 call as rank 3 of 8 (device-side shard) for a launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
gal = torch.from_numpy(synthetic.uniform_points(43, 1000000, 1000.0)).cuda()
ptcl = torch.from_numpy(synthetic.uniform_points(44, 100000000, 1000.0)).cuda()
rp = np.logspace(-1, np.log10(30), 15)
_lib.set_shard(3, 8)
for _ in range(2):
    hb.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0)
    print(_lib.last_stats)
