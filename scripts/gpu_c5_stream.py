"""config-5 call: wall time with the library stream and with a caller-provided stream (as bench.py uses)."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
gal = torch.from_numpy(synthetic.uniform_points(43, 1000000, 1000.0)).cuda()
ptcl = torch.from_numpy(synthetic.uniform_points(44, 100000000, 1000.0)).cuda()
rp = np.logspace(-1, np.log10(30), 15)
def run(tag):
    for i in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hb.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = _lib.last_stats
        print(tag, i, "wall %.2f total %.2f count %.2f mesh %.2f h2d %.3f" % (dt * 1e3, st["ms_total"], st["ms_count"], st["ms_mesh"], st["ms_h2d"]), flush=True)
run("lib-stream")
stream = torch.cuda.Stream()
_lib.check(_lib.load().htb_set_stream(ctypes.c_void_p(stream.cuda_stream)))
with torch.cuda.stream(stream):
    run("user-stream")
_lib.collect_stats = False
with torch.cuda.stream(stream):
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hb.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0)
        torch.cuda.synchronize(); print("user-stream no-stats wall %.2f" % ((time.perf_counter() - t0) * 1e3), flush=True)
