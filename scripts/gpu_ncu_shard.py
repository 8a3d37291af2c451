import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
rb = synthetic.config_rbins()
for world, rank in ((1, 0), (8, 3)):
    _lib.set_shard(rank, world)
    hb.npairs_3d(ran, ran, rb, period=250.0)
    print(world, rank, _lib.last_stats["ms_count"], _lib.last_stats["pairs_evaluated"])
