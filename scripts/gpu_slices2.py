import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
rb = synthetic.config_rbins()
def run(a, b):
    best = 1e9
    for _ in range(3):
        hb.npairs_3d(a, b, rb, period=250.0)
        best = min(best, _lib.last_stats["ms_count"])
    return round(best, 3), _lib.last_stats["tiles"], _lib.last_stats["tiles_redone"], _lib.last_stats["pairs_evaluated"]
for world, rank in ((1, 0), (8, 3)):
    _lib.set_shard(rank, world)
    for sym in ("0", "1"):
        for ms in ("1", "2", "4", "8", "16"):
            os.environ["HTB_MAXSLICES"] = ms
            os.environ["HTB_ITEMS_PER_WARP"] = "1024"
            if sym == "1": os.environ["HTB_NO_SYM"] = "1"
            else: os.environ.pop("HTB_NO_SYM", None)
            print(world, rank, "nosym", sym, "maxslices", ms, "RR", run(ran, ran), flush=True)
