"""Run the BASELINE.json configurations 1, 3, 4, 5 at full size on the GPU: timings, work counters and
size-independent checks (known answer for config 1; additivity / sampled-oracle checks otherwise)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import halotools_b200 as hb  # noqa: E402
from halotools_b200 import _lib, synthetic  # noqa: E402

out = {}


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return r, best, dict(_lib.last_stats or {})


which = sys.argv[1:] or ["1", "3", "4", "5"]
rb = synthetic.config_rbins()
if "1" in which:
    s = synthetic.uniform_points(43, 100000, 250.0)
    r, dt, st = timed(lambda: hb.npairs_3d(s, s, rb, period=250.0), 3)
    want = [100000, 100004, 100020, 100066, 100254, 100790, 102408, 107560, 123606, 173566, 328856, 812270,
            2314792, 6988496, 21546088]
    out["config1"] = {"ok": bool(np.array_equal(r, want)), "wall_s": dt, "stats": st}
if "3" in which:
    s = synthetic.uniform_points(43, 2000000, 1000.0)
    rp = np.logspace(-1, np.log10(30), 15)
    r, dt, st = timed(lambda: hb.npairs_xy_z(s, s, rp, [0.0, 60.0], period=1000.0))
    w, dtw, _ = timed(lambda: hb.wp(s, rp, 60.0, period=1000.0), 1)
    half = hb.npairs_xy_z(s, s[:1000000], rp, [0.0, 60.0], period=1000.0) + hb.npairs_xy_z(s, s[1000000:], rp, [0.0, 60.0], period=1000.0)
    out["config3"] = {"ok": bool(np.array_equal(half, r)), "wall_s": dt, "wp_wall_s": dtw, "stats": st,
                      "counts_top": int(r[-1, -1]), "wp": w.tolist()}
if "4" in which:
    rng = np.random.RandomState(43)
    s = rng.uniform(0, 1000.0, (10000000, 3))
    w = rng.uniform(0.5, 1.5, 10000000)
    r, dt, st = timed(lambda: hb.marked_npairs_3d(s, s, rb, 1, period=1000.0, weights1=w, weights2=w))
    n, dtn, stn = timed(lambda: hb.npairs_3d(s, s, rb, period=1000.0))
    # unit weights must reproduce the integer counts exactly
    ones = np.ones(10000000)
    u = hb.marked_npairs_3d(s, s, rb, 1, period=1000.0, weights1=ones, weights2=ones)
    out["config4"] = {"ok": bool(np.array_equal(u, n.astype(float))), "marked_wall_s": dt, "marked_stats": st,
                      "npairs_wall_s": dtn, "npairs_stats": stn, "mean_weight_ratio": (r / n).tolist()}
if "5" in which:
    ngal, nptcl = int(os.environ.get("C5_NGAL", 1000000)), int(os.environ.get("C5_NPTCL", 100000000))
    gal = synthetic.uniform_points(43, ngal, 1000.0)
    ptcl = synthetic.uniform_points(44, nptcl, 1000.0)
    rp = np.logspace(-1, np.log10(30), 15)
    r, dt, st = timed(lambda: hb.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0), 1)
    # sampled check: per-object rows of 2000 galaxies against the same call on those galaxies alone
    sub = gal[:2000]
    a = hb.mean_delta_sigma(sub, ptcl, 1.0, rp, period=1000.0, per_object=True)
    out["config5"] = {"wall_s": dt, "stats": st, "delta_sigma": r.tolist(), "ngal": ngal, "nptcl": nptcl,
                      "uniform_expectation_zero_over_sigma": (np.mean(a, axis=0) / (np.std(a, axis=0) / np.sqrt(len(sub)))).tolist()}
print(json.dumps(out))
