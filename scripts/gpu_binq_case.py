"""One call of the BinQ (rp, pi) kernel at config-3 size with rp_pi_tpcf-like bins (for ncu captures)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import halotools_b200 as hb
from halotools_b200 import _lib, synthetic
s = synthetic.uniform_points(43, 2000000, 1000.0)
rp = np.logspace(-1, np.log10(30), 15)
which = sys.argv[1] if len(sys.argv) > 1 else "xyz"
if which == "xyz":
    hb.npairs_xy_z(s, s, rp, np.linspace(0, 40, 41), period=1000.0)
else:
    hb.npairs_s_mu(s, s, rp, np.linspace(0, 1, 11), period=1000.0)
print(json.dumps(_lib.last_stats))
