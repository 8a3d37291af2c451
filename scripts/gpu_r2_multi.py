"""Multi-GPU check under torchrun (one rank per GPU, NCCL): the sharded statistics - device-side all-reduce of the count
tables, 1/world uploads + all-gather, work-balanced cell ranges - against the same calls with sharding switched off on
every rank.  Integer-count statistics must be IDENTICAL; delta-sigma within 1e-12."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import halotools_b200 as hb  # noqa: E402
from halotools_b200 import _lib, distributed, synthetic  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_lib.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

gal = synthetic.fakesim_zheng07_mock(140, 250.0, seed=43)
ran = synthetic.uniform_points(44, 1500000, 250.0)
rb = synthetic.config_rbins()
rp = np.logspace(-1, np.log10(20.0), 12)
g2 = synthetic.uniform_points(45, 200000, 400.0)
p2 = synthetic.uniform_points(46, 3000000, 400.0)


def run_all():
    out = {}
    out["tpcf_ls"] = hb.tpcf(gal, rb, randoms=ran, period=250.0, estimator="Landy-Szalay")
    out["tpcf_dev"] = hb.tpcf(torch.from_numpy(gal).cuda(), rb, randoms=torch.from_numpy(ran).cuda(), period=250.0,
                              estimator="Landy-Szalay")
    out["tpcf_cross"] = np.concatenate(hb.tpcf(gal[::2], rb, sample2=gal[1::2], period=250.0))
    out["wp"] = hb.wp(gal, rp, 40.0, period=250.0)
    out["rp_pi"] = hb.rp_pi_tpcf(gal, rp, np.linspace(0, 40, 9), randoms=ran[:400000], period=250.0, estimator="Natural").ravel()
    out["n3d"] = hb.npairs_3d(gal, ran, rb, period=250.0).astype(float)
    out["marked"] = hb.marked_npairs_3d(gal, gal, rb, 1, period=250.0, weights1=np.linspace(0.5, 1.5, len(gal)),
                                        weights2=np.linspace(0.5, 1.5, len(gal)))
    out["ds"] = hb.mean_delta_sigma(g2, p2, 1.0, rp, period=400.0)
    out["ds_rows"] = hb.mean_delta_sigma(g2[:5000], p2, 1.0, rp, period=400.0, per_object=True).ravel()
    return out


distributed.enable()
sharded = run_all()
distributed.disable()
alone = run_all()
ok = True
for k in alone:
    a, b = np.asarray(sharded[k]), np.asarray(alone[k])
    if k in ("ds", "ds_rows", "marked"):
        good = np.allclose(a, b, rtol=1e-11, atol=1e-12 * np.max(np.abs(b)))
    else:
        good = np.array_equal(a, b)
    if not good:
        ok = False
        print("rank %d MISMATCH %s max|diff| %.3g" % (rank, k, float(np.max(np.abs(a - b)))), flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("multi-GPU check world=%d: %s" % (world, "OK" if int(t.item()) == 1 else "FAILED"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
