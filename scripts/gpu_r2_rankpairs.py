"""Per-rank work counters of the 8-GPU RR shard (one GPU runs every rank's shard): predicted vs evaluated pairs, tiles, time."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
gal = torch.from_numpy(synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)).cuda()
rb = synthetic.config_rbins()
world = 8
for name, a, b in (("RR", ran, ran), ("DR", gal, ran), ("DD", gal, gal)):
    rows = []
    for r in range(world):
        _lib.set_shard(r, world)
        hb.npairs_3d(a, b, rb, period=250.0)
        hb.npairs_3d(a, b, rb, period=250.0)
        st = _lib.last_stats
        rows.append((r, round(st["ms_count"], 3), round(st["ms_mesh"], 3), st["tiles"], "%.4g" % st["pairs_evaluated"], "%.4g" % st["pairs_reference"], st["tiles_redone"]))
    print(name, "rank, ms_count, ms_mesh, tiles, evaluated, reference, redone")
    for row in rows:
        print("  ", row)
_lib.set_shard(0, 1)
