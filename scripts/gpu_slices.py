"""Sweep of the tile-slicing knobs on one rank's shard of the bench step (world 8, rank 3)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
gal = torch.from_numpy(synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)).cuda()
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
rb = synthetic.config_rbins()
def run(a, b):
    best = 1e9
    for _ in range(3):
        hb.npairs_3d(a, b, rb, period=250.0)
        best = min(best, _lib.last_stats["ms_count"])
    return round(best, 3), _lib.last_stats["tiles"], _lib.last_stats["tiles_redone"], _lib.last_stats["pairs_evaluated"]
for world, rank in ((1, 0), (8, 3), (8, 0)):
    _lib.set_shard(rank, world)
    for ms in ("1", "2", "3", "4", "8", "16"):
        for ipw in ("12", "32"):
            os.environ["HTB_MAXSLICES"] = ms
            os.environ["HTB_ITEMS_PER_WARP"] = ipw
            print(world, rank, "maxslices", ms, "ipw", ipw, "DD", run(gal, gal), "DR", run(gal, ran), "RR", run(ran, ran), flush=True)
