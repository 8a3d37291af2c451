"""Timings of the SURVEY 8(f) counters at config-3-like sizes (GPU wall time of the public call, count-kernel time,
evaluated pairs) next to the CPU oracle port on a subsample (1 thread; the reference's jackknife engine does
O(N_samples) work per pair, the port is timed at 20000 points)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
from oracle import oracle
out = {}
def run(name, fn, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = dict(_lib.last_stats)
    out[name] = {"wall_ms": best * 1e3, "ms_count": st["ms_count"], "ms_mesh": st["ms_mesh"], "ms_h2d": st["ms_h2d"],
                 "pairs_evaluated": st["pairs_evaluated"], "path": st["path"]}
    print(name, json.dumps(out[name]), flush=True)
    return r
def cpu(name, fn):
    t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
    out[name] = {"cpu_port_s": dt}
    print(name, dt, flush=True)
L = 1000.0
s = synthetic.uniform_points(43, 2000000, L)
rng = np.random.RandomState(5)
w = rng.uniform(0.5, 1.5, len(s))
rp = np.logspace(-1, np.log10(30), 15)
rb = synthetic.config_rbins()
run("npairs_projected 2e6, pi_max 60", lambda: hb.npairs_projected(s, s, rp, 60.0, period=L))
run("npairs_per_object_3d 2e6, 15 rbins to 20", lambda: hb.npairs_per_object_3d(s, s, rb, period=L))
run("marked_npairs_xy_z 2e6, id 1, 15 x 21 bins", lambda: hb.marked_npairs_xy_z(s, s, rp, np.linspace(0, 40, 21), period=L, weights1=w, weights2=w, weight_func_id=1))
tags, nsub = hb.catalog_analysis_helpers.cuboid_subvolume_labels(s, 5, L) if hasattr(hb, "catalog_analysis_helpers") else (None, None)
if tags is None:
    from halotools_b200.catalog_analysis_helpers import cuboid_subvolume_labels
    tags, nsub = cuboid_subvolume_labels(s, 5, L)
run("npairs_jackknife_3d 2e6, 125 sub-volumes", lambda: hb.npairs_jackknife_3d(s, s, rb, tags, tags, nsub, period=L, weights1=w, weights2=w))
run("npairs_jackknife_xy_z 2e6, 125 sub-volumes, 15 x 2", lambda: hb.npairs_jackknife_xy_z(s, s, rp, [0.0, 60.0], tags, tags, nsub, period=L))
run("npairs_jackknife_xy_z 2e6, 125 sub-volumes, 15 x 41 (rows in global memory)", lambda: hb.npairs_jackknife_xy_z(s, s, rp, np.linspace(0, 40, 41), tags, tags, nsub, period=L))
def wall(name, fn, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out[name] = {"wall_ms": best * 1e3}
    print(name, best * 1e3, flush=True)
wall("statistic wp 2e6, pi_max 60 (analytic randoms)", lambda: hb.wp(s, rp, 60.0, period=L))
wall("statistic rp_pi_tpcf 2e6, 15 x 41 bins (analytic randoms)", lambda: hb.rp_pi_tpcf(s, rp, np.linspace(0, 40, 41), period=L))
wall("statistic s_mu_tpcf 2e6, 15 x 11 bins (analytic randoms)", lambda: hb.s_mu_tpcf(s, rp, np.linspace(0, 1, 11), period=L))
hid = rng.randint(0, 400000, len(s))
wall("statistic tpcf_one_two_halo_decomp 2e6 (analytic randoms)", lambda: hb.tpcf_one_two_halo_decomp(s, hid, rb, period=L))
gal = synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)
ran = synthetic.uniform_points(44, 2000000, 250.0)
t0 = time.perf_counter(); xi, cov = hb.tpcf_jackknife(gal, ran, rb, Nsub=5, period=250.0, estimator="Landy-Szalay"); dt = time.perf_counter() - t0
out["tpcf_jackknife zheng07 mock (%d) + 2e6 randoms, Nsub 5, LS" % len(gal)] = {"wall_ms": dt * 1e3, "xi0": float(xi[0]), "cov_diag0": float(cov[0, 0])}
print("tpcf_jackknife", dt, flush=True)
g2 = np.ascontiguousarray(s[:200000, :2]); p2 = synthetic.uniform_points(44, 10000000, L)[:, :2].copy(); m2 = rng.uniform(0, 2, len(p2))
run("weighted_npairs_xy 2e5 x 1e7", lambda: hb.weighted_npairs_xy(g2, p2, m2, rp, period=L))
run("weighted_npairs_per_object_xy 2e5 x 1e7", lambda: hb.weighted_npairs_per_object_xy(g2, p2, m2, rp, period=L))
sm = s[:20000] * 0.1           # same number density in a (100)^3 box
tg, ns = __import__("halotools_b200.catalog_analysis_helpers", fromlist=["x"]).cuboid_subvolume_labels(sm, 5, 100.0)
cpu("oracle port npairs_jackknife_3d 2e4 points, 125 sub-volumes (1 thread)", lambda: oracle.npairs_jackknife_3d(sm, sm, np.logspace(-1, 0.8, 15), tg, tg, ns, period=100.0))
cpu("oracle port npairs_per_object_3d 2e5 points (1 thread)", lambda: oracle.npairs_per_object_3d(s[:200000] * 0.464, s[:200000] * 0.464, rb, period=464.0))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r01_8f_counters.json"), "w"), indent=1)
