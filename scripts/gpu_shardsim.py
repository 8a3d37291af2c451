"""Predict multi-GPU strong scaling on ONE GPU: for world = 2, 4, 8 run every rank's shard of the bench
step (DD, DR, RR of configs[1]) or of config 5 one after the other and report per-rank times; the
max over ranks is what a real N-GPU step would take (plus the all-reduce)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import halotools_b200 as hb  # noqa: E402
from halotools_b200 import _lib, synthetic, distributed  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "tpcf"
worlds = [int(w) for w in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4", "8"])]
BALANCED = os.environ.get("SHARD_MODE", "balanced") == "balanced"
fake = {"rank": 0, "world": 1}
if not BALANCED:
    # the reference's equal-cell-count ranges, applied on the host
    distributed._rank_world = lambda: (fake["rank"], fake["world"])
    distributed.allreduce_sum = lambda a: a
    distributed._state["enabled"] = True

if which == "tpcf":
    gal = torch.from_numpy(synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)).cuda()
    ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
    rb = synthetic.config_rbins()

    def step():
        ms = []
        for a, b in ((gal, gal), (gal, ran), (ran, ran)):
            hb.npairs_3d(a, b, rb, period=250.0)
            ms.append((_lib.last_stats["ms_total"], _lib.last_stats["ms_count"], _lib.last_stats["ms_mesh"]))
        return ms
else:
    ngal, nptcl = int(os.environ.get("C5_NGAL", 1000000)), int(os.environ.get("C5_NPTCL", 100000000))
    gal = torch.from_numpy(synthetic.uniform_points(43, ngal, 1000.0)).cuda()
    ptcl = torch.from_numpy(synthetic.uniform_points(44, nptcl, 1000.0)).cuda()
    rp = np.logspace(-1, np.log10(30), 15)

    def step():
        hb.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0)
        st = _lib.last_stats
        return [(st["ms_total"], st["ms_count"], st["ms_mesh"])]

out = {}
for world in worlds:
    per_rank = []
    for r in range(world):
        fake["rank"], fake["world"] = r, world
        if BALANCED:
            _lib.set_shard(r, world)
        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms = step()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        per_rank.append({"wall_ms": wall, "calls_total_count_mesh_ms": ms})
    walls = [p["wall_ms"] for p in per_rank]
    out[str(world)] = {"max_wall_ms": max(walls), "mean_wall_ms": float(np.mean(walls)), "per_rank": per_rank}
    sys.stderr.write("world %d: max %.2f mean %.2f\n" % (world, max(walls), float(np.mean(walls))))
base = out[str(worlds[0])]["max_wall_ms"] * worlds[0]
for w in worlds:
    out[str(w)]["predicted_efficiency"] = base / (w * out[str(w)]["max_wall_ms"])
print(json.dumps(out))
