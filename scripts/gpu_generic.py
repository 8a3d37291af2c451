"""Timings of the kernels behind the remaining front-ends at config-3/4 sizes (generic vs fast paths)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
out = {}
def run(name, fn, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = dict(_lib.last_stats)
    out[name] = {"wall_ms": best * 1e3, "ms_count": st["ms_count"], "ms_mesh": st["ms_mesh"], "ms_h2d": st["ms_h2d"],
                 "pairs_evaluated": st["pairs_evaluated"], "path": st["path"], "tiles": st["tiles"],
                 "gpairs_per_s": st["pairs_evaluated"] / st["ms_count"] / 1e6}
    print(name, json.dumps(out[name]), flush=True)
    return r
rb = synthetic.config_rbins()
s2m = synthetic.uniform_points(43, 2000000, 1000.0)
rp = np.logspace(-1, np.log10(30), 15)
run("c3 npairs_xy_z wp (2 pi edges, fast)", lambda: hb.npairs_xy_z(s2m, s2m, rp, [0.0, 60.0], period=1000.0))
run("c3 npairs_xy_z 12 pi edges (BinQ)", lambda: hb.npairs_xy_z(s2m, s2m, rp, np.linspace(0, 60, 12), period=1000.0))
run("c3 npairs_xy_z 41 pi edges (BinQ)", lambda: hb.npairs_xy_z(s2m, s2m, rp, np.linspace(0, 40, 41), period=1000.0))
run("c3 npairs_3d 40 rbins (BinQ)", lambda: hb.npairs_3d(s2m, s2m, np.linspace(0.1, 30, 40), period=1000.0))
run("c3 npairs_s_mu 15 x 11 (BinQ)", lambda: hb.npairs_s_mu(s2m, s2m, rp, np.linspace(0, 1, 11), period=1000.0))
run("c3 npairs_3d (fast)", lambda: hb.npairs_3d(s2m, s2m, rp, period=1000.0))
rng = np.random.RandomState(43)
s = rng.uniform(0, 1000.0, (10000000, 3)); w = rng.uniform(0.5, 1.5, 10000000)
run("c4 marked id 1 (fast)", lambda: hb.marked_npairs_3d(s, s, rb, 1, period=1000.0, weights1=w, weights2=w))
run("c4 marked id 2 (generic)", lambda: hb.marked_npairs_3d(s, s, rb, 2, period=1000.0, weights1=w, weights2=w))
run("c4 npairs_3d (fast)", lambda: hb.npairs_3d(s, s, rb, period=1000.0))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "generic.json"), "w"), indent=1)
