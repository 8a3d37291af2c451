"""DD / DR / RR count times of the bench for small slice caps (unsharded and shard 3 of 8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
gal = torch.from_numpy(synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)).cuda()
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
rb = synthetic.config_rbins()
for world, rank in ((1, 0), (2, 1), (4, 1), (8, 3)):
    _lib.set_shard(rank, world)
    for ms in (None,):
        if ms is None: os.environ.pop("HTB_ITEMS_PER_WARP", None)
        else: os.environ["HTB_ITEMS_PER_WARP"] = ms
        row = []
        for a, b in ((gal, gal), (gal, ran), (ran, ran)):
            hb.npairs_3d(a, b, rb, period=250.0)
            hb.npairs_3d(a, b, rb, period=250.0)
            row.append((round(_lib.last_stats["ms_count"], 3), _lib.last_stats["tiles"], "%.3g" % _lib.last_stats["pairs_evaluated"]))
        print("world", world, "items_per_warp", ms, "DD DR RR (count ms, tiles, pairs)", row, flush=True)
_lib.set_shard(0, 1)
