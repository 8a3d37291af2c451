"""Config 1 (1e5 points, the MCMC-sized call): wall time per call from host / device-resident input, and the statistic tpcf
with analytic randoms (one count + estimator)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
s = synthetic.uniform_points(43, 100000, 250.0)
d = torch.from_numpy(s).cuda()
rb = synthetic.config_rbins()


def wall(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("npairs_3d host input   : %.3f ms/call" % wall(lambda: hb.npairs_3d(s, s, rb, period=250.0)), _lib.last_stats)
print("npairs_3d device input : %.3f ms/call" % wall(lambda: hb.npairs_3d(d, d, rb, period=250.0)), {k: _lib.last_stats[k] for k in ("ms_mesh", "ms_count", "ms_total", "kernel_launches", "tiles")})
_lib.collect_stats = False
print("npairs_3d device input, no stats : %.3f ms/call" % wall(lambda: hb.npairs_3d(d, d, rb, period=250.0)))
_lib.collect_stats = True
print("tpcf (analytic randoms) host   : %.3f ms/call" % wall(lambda: hb.tpcf(s, rb, period=250.0)))
print("tpcf (analytic randoms) device : %.3f ms/call" % wall(lambda: hb.tpcf(d, rb, period=250.0)))
print("wp device (2e5 pts)            : %.3f ms/call" % wall(lambda: hb.wp(d, np.logspace(-1, 1.2, 12), 40.0, period=250.0)))
