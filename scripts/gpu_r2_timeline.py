"""Where one rank's tpcf step spends its time: events behind every set-up / count of hb.tpcf (shard `rank` of 8 on one GPU)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
from halotools_b200.two_point_clustering import _device
gal = torch.from_numpy(synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)).cuda()
ran = torch.from_numpy(synthetic.uniform_points(44, 5000000, 250.0)).cuda()
rb = synthetic.config_rbins()
stream = _lib.engine_stream()
import time
HOST = []
_orig = _lib.run_engine


def _timed_engine(name, *a, **k):
    t0 = time.perf_counter()
    r = _orig(name, *a, **k)
    HOST.append((name, t0, time.perf_counter()))
    return r


_lib.run_engine = _timed_engine
for world, rank in ((1, 0), (8, 3)):
    _lib.set_shard(rank, world)
    for rep in range(3):
        _device.TIMELINE = [] if rep == 2 else None
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        del HOST[:]
        th0 = time.perf_counter()
        try:
            hb.tpcf(gal, rb, randoms=ran, period=250.0, estimator="Landy-Szalay")
        except ValueError:
            pass
        e1.record(stream)
        torch.cuda.synchronize()
    print("world %d rank %d: step %.2f ms" % (world, rank, e0.elapsed_time(e1)))
    print("   host: engine calls (start, end) ms after the step began:", [(n.replace("htb_", ""), round((a - th0) * 1e3, 2), round((b - th0) * 1e3, 2)) for n, a, b in HOST])
    for what, si, ev in _device.TIMELINE:
        print("   %-8s stream %d done at %.2f ms" % (what, si, e0.elapsed_time(ev)))
    st = _lib.async_kernel_stamps()[-3:]
    t0 = min(a for a, b in st if a)
    print("   count kernels, first warp in -> last warp out (ms after the first one starts):",
          [(round((a - t0) * 1e-6, 2), round((b - t0) * 1e-6, 2)) for a, b in st])
    print("   count kernel brackets (ms):", [round(t, 2) for t in _lib.async_count_times()][-3:])
_device.TIMELINE = None
_lib.set_shard(0, 1)
