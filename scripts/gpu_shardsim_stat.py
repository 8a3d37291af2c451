"""Every rank's shard of the N-GPU tpcf step (hb.tpcf: device-resident samples, K3 on the device) run one after the other on
ONE GPU: device-timed per rank, max over ranks = what the real N-GPU step takes apart from the all-reduce.
usage: gpu_shardsim_stat.py <worlds, e.g. 1,8>   (env HTB_ONE_STREAM, HTB_EARLY_EXIT, HTB_TAIL_EIGHTHS select variants)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import halotools_b200 as hb  # noqa: E402
from halotools_b200 import _lib, synthetic  # noqa: E402

worlds = [int(w) for w in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "8"])]
gal = torch.from_numpy(synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)).cuda()
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
rb = synthetic.config_rbins()
stream = _lib.engine_stream()


def step():
    # (a rank's partial counts give a meaningless xi: a zero RR bin must not raise here)
    try:
        return hb.tpcf(gal, rb, randoms=ran, period=250.0, estimator="Landy-Szalay")
    except ValueError:
        return None


out = {}
for world in worlds:
    per_rank = []
    for r in range(world):
        _lib.set_shard(r, world)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        per_rank.append(e0.elapsed_time(e1) / 3)
    out[str(world)] = {"per_rank_ms": per_rank, "max_ms": max(per_rank)}
_lib.set_shard(0, 1)
base = out[str(worlds[0])]["max_ms"] * worlds[0]
for w in worlds:
    out[str(w)]["predicted_efficiency"] = base / (w * out[str(w)]["max_ms"])
print(json.dumps(out))
