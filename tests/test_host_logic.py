"""CPU tests (no GPU) of the host layer: argument validation with the reference's error types and
message fragments (the strings the reference's own tests assert on), mesh geometry against the
oracle's restatement, estimators, the C-ABI library (loads, exports every declared symbol, no compute
call), and the multi-rank sharding logic over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import halotools_b200 as hb
from halotools_b200 import _lib, distributed
from halotools_b200.custom_exceptions import HalotoolsError
from halotools_b200.pair_counters.mesh_helpers import double_mesh_geometry, _cell1_parallelization_indices
from halotools_b200.pair_counters.npairs_3d import _npairs_3d_process_args
from halotools_b200.pair_counters.marked_npairs_3d import _marked_npairs_process_weights, _func_signature_int_from_wfunc
from halotools_b200.two_point_clustering.tpcf_estimators import _TP_estimator, _TP_estimator_crossx
from oracle import oracle
from oracle.mesh import DoubleMesh
from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S = cases.pts(43, 100)


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.htb_abi_version() == 1
    header = open(os.path.join(ROOT, "include", "halotools_b200.h")).read()
    declared = set(re.findall(r"^(?:const char \*|int\s+)(htb_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    assert ctypes.sizeof(_lib.MeshGeom) == 2 * 4 + 3 * 3 * 4 + 4 + 4 * 3 * 8
    assert isinstance(lib.htb_device_count(), int)


def test_no_cpu_fallback_without_gpu():
    if _lib.load().htb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.npairs_3d(S, S, [0.1, 0.2], period=1.0)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.mean_delta_sigma(S, S, 1.0, [0.1, 0.2], period=1.0)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "halotools_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)


@pytest.mark.parametrize("bad", [2.5, "Cuba Gooding Jr.", np.int64(2)])
def test_num_threads_validation(bad):
    with pytest.raises(ValueError, match="Input ``num_threads`` argument must be an integer or the string 'max'"):
        hb.npairs_3d(S, S, [0.1, 0.2], period=1.0, num_threads=bad)
    with pytest.raises(ValueError, match="num_threads"):
        hb.npairs_xy_z(S, S, [0.1, 0.2], [0.1, 0.2], period=1.0, num_threads=bad)


def test_rbins_and_period_validation():
    with pytest.raises(ValueError, match="Input ``rbins`` must be a monotonically increasing 1D array with at least two entries"):
        hb.npairs_3d(S, S, [0.1], period=1.0)
    with pytest.raises(ValueError, match="rbins"):
        hb.npairs_3d(S, S, [0.1, 0.3, 0.2], period=1.0)
    with pytest.raises(ValueError, match="Input ``period`` must be a bounded positive number in all dimensions"):
        hb.npairs_3d(S, S, [0.1, 0.2], period=[1.0, 1.0, np.inf])
    with pytest.raises(ValueError, match="period"):
        hb.npairs_3d(S, S, [0.1, 0.2], period=-1.0)
    with pytest.raises(ValueError, match="The maximum length over which you search for pairs of points"):
        hb.npairs_3d(S, S, [0.1, 0.4], period=1.0)
    with pytest.raises(ValueError, match="rp_bins"):
        hb.npairs_xy_z(S, S, [0.1], [0.1, 0.2], period=1.0)
    with pytest.raises(ValueError, match="pi_bins"):
        hb.npairs_xy_z(S, S, [0.1, 0.2], [0.2, 0.1, 0.3], period=1.0)
    with pytest.raises(ValueError, match="mu_bins"):
        hb.npairs_s_mu(S, S, [0.1, 0.2], [0.5], period=1.0)
    with pytest.raises(ValueError, match="must be a length-3 sequence"):
        hb.npairs_3d(S, S, [0.1, 0.2], period=1.0, approx_cell1_size=[0.1, 0.1])


def test_process_args_defaults():
    r = _npairs_3d_process_args(S, S, [0.1, 0.25], None, 1, None, None)
    period, pbcs = r[7], r[9]
    assert pbcs is False and np.allclose(period, max(np.ptp(S), 0.75))
    assert r[10] == [0.25, 0.25, 0.25] and r[11] == [0.25, 0.25, 0.25]
    assert min(np.min(c) for c in r[:6]) == 0.0
    r = _npairs_3d_process_args(S, S, [0.1, 0.25], 2.0, "max", 0.3, [0.1, 0.2, 0.3])
    assert np.array_equal(r[7], [2.0, 2.0, 2.0]) and r[9] is True and r[10] == [0.3, 0.3, 0.3]


def test_weights_validation():
    for wid, nw in cases.NUM_WEIGHTS.items():
        assert _func_signature_int_from_wfunc(wid) == nw
        w1, w2 = _marked_npairs_process_weights(S, S, np.ones((100, nw)), None, wid)
        assert w1.shape == (100, nw) and w2.shape == (100, nw) and np.all(w2 == 1)
        with pytest.raises(HalotoolsError):
            _marked_npairs_process_weights(S, S, np.ones((100, nw + 1)), None, wid)
    with pytest.raises(HalotoolsError, match="does not have the correct length"):
        _marked_npairs_process_weights(S, S, np.ones(99), None, 1)
    with pytest.raises(HalotoolsError, match="is not recognized"):
        _func_signature_int_from_wfunc(18)
    with pytest.raises(ValueError, match="must be an integer ID"):
        _func_signature_int_from_wfunc(1.0)
    with pytest.raises(HalotoolsError, match="1-D or 2-D array"):
        _marked_npairs_process_weights(S, S, np.ones((100, 1, 1)), None, 1)


def test_clustering_level_validation():
    rb = np.logspace(-2, -1, 5)
    with pytest.raises(TypeError, match="Input sample of points must be a Numpy ndarray of shape"):
        hb.tpcf(np.ones((10, 2)), rb, period=1.0)
    with pytest.raises(TypeError, match="strictly positive"):
        hb.tpcf(S, [0.0, 0.1, 0.2], period=1.0)
    with pytest.raises(ValueError, match="If no PBCs are specified, randoms must be provided"):
        hb.tpcf(S, rb, period=None)
    with pytest.raises(ValueError, match="is not in the list of available estimators"):
        hb.tpcf(S, rb, period=1.0, estimator="Jose Canseco")
    with pytest.raises(ValueError, match="does not permit you to look for pairs"):
        hb.tpcf(S, [0.1, 0.4], period=1.0)
    with pytest.raises(HalotoolsError, match="You must either provide both"):
        hb.tpcf(S, rb, period=1.0, RR_precomputed=np.ones(4))
    with pytest.raises(HalotoolsError, match="must match length"):
        hb.tpcf(S, rb, period=1.0, randoms=S, RR_precomputed=np.ones(3), NR_precomputed=100)
    with pytest.raises(ValueError, match="`normalize_by` parameter not recognized"):
        hb.marked_tpcf(S, rb, period=1.0, normalize_by="Arnold Schwarzenegger")
    with pytest.raises(HalotoolsError, match="`marks1` must have same length as `sample1`"):
        hb.marked_tpcf(S, rb, period=1.0, marks1=np.ones(7))
    with pytest.raises(ValueError, match="there are values in the x-dimension"):
        hb.mean_delta_sigma(S + np.array([1.0, 0, 0]), S, 1.0, rb, period=1.0)
    with pytest.raises(ValueError, match="your input data has negative values"):
        hb.mean_delta_sigma(S - 0.5, S, 1.0, rb, period=1.0)


@pytest.mark.parametrize("n1", [1, 7, 60])
@pytest.mark.parametrize("L,search", [(1.0, 0.3), (250.0, 20.0), (1000.0, 30.0), (3.0, 1.0)])
@pytest.mark.parametrize("a1,a2", [(None, None), (0.07, 0.013), (0.5, 0.25)])
def test_mesh_geometry_matches_oracle_mesh(n1, L, search, a1, a2):
    pts1 = cases.pts(1, n1, L)
    ac1 = [search] * 3 if a1 is None else [a1 * L] * 3
    ac2 = [search] * 3 if a2 is None else [a2 * L] * 3
    geom = double_mesh_geometry(3, ac1, ac2, [search] * 3, [L] * 3, True)
    dm = DoubleMesh([pts1[:, d] for d in range(3)], [pts1[:, d] for d in range(3)], ac1, ac2, [search] * 3, [L] * 3, True)
    assert geom.ndivs1 == dm.mesh1.num_divs and geom.ndivs2 == dm.mesh2.num_divs
    assert geom.cell1_size == dm.mesh1.cell_size and geom.cell2_size == dm.mesh2.cell_size
    assert geom.cover == dm.cover
    for d in range(3):
        assert geom.ndivs2[d] % geom.ndivs1[d] == 0 and geom.ndivs1[d] >= 3 and geom.ndivs2[d] <= 50
        assert geom.cell1_size[d] >= search * (1 - 1e-12)


def test_reference_mesh_configs():
    # SURVEY.md 8a row a2
    assert double_mesh_geometry(3, [20.0] * 3, [20.0] * 3, [20.0] * 3, [250.0] * 3, True).ndivs1 == [12, 12, 12]
    assert double_mesh_geometry(3, [30.0, 30.0, 60.0], [30.0, 30.0, 60.0], [30.0, 30.0, 60.0], [1000.0] * 3, True).ndivs1 == [33, 33, 16]
    assert double_mesh_geometry(3, [20.0] * 3, [20.0] * 3, [20.0] * 3, [1000.0] * 3, True).ndivs1 == [50, 50, 50]
    assert double_mesh_geometry(2, [30.0] * 2, [30.0] * 2, [30.0] * 2, [1000.0] * 2, True).ndivs1 == [33, 33]


def test_cell1_parallelization_indices():
    assert _cell1_parallelization_indices(10, 1) == (1, [(0, 10)])
    n, t = _cell1_parallelization_indices(3, 5)
    assert n == 3 and t == [(0, 1), (1, 2), (2, 3)]
    n, t = _cell1_parallelization_indices(10, 3)
    assert t == [(0, 4), (4, 7), (7, 10)]
    assert distributed.split_cells(10, 3) == [(0, 4), (4, 7), (7, 10)]
    work = np.array([0, 0, 10, 0, 0, 10, 0, 0, 10, 10.0])
    parts = distributed.split_cells(10, 2, work)
    assert parts[0][0] == 0 and parts[-1][1] == 10 and parts[0][1] == parts[1][0]
    assert abs(work[parts[0][0]:parts[0][1]].sum() - 20.0) <= 10.0


def test_estimators():
    DD, DR, RR = np.array([10.0, 20.0]), np.array([5.0, 8.0]), np.array([4.0, 2.0])
    assert np.allclose(_TP_estimator(DD, DR, RR, 10, 10, 20, 20, "Natural"), DD / RR * 4 - 1)
    assert np.allclose(_TP_estimator(DD, DR, RR, 10, 10, 20, 20, "Landy-Szalay"), DD / RR * 4 - 2 * DR / RR * 2 + 1)
    assert np.allclose(_TP_estimator(DD, DR, RR, 10, 10, 20, 20, "Hamilton"), DD * RR / DR ** 2 - 1)
    assert np.allclose(_TP_estimator_crossx(DD, DR, DR, RR, 10, 10, 20, 20, "Hamilton"), DD * RR / DR ** 2 - 1)
    with pytest.raises(ValueError, match="zero RR pairs"):
        _TP_estimator(DD, DR, np.array([0.0, 1.0]), 10, 10, 20, 20, "Natural")
    with pytest.raises(ValueError, match="zero DR pairs"):
        _TP_estimator(DD, np.array([0.0, 1.0]), RR, 10, 10, 20, 20, "Hamilton")
    with pytest.raises(ValueError, match="not supported for cross-correlations"):
        _TP_estimator_crossx(DD, DR, DR, RR, 10, 10, 20, 20, "Hewett")


GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from halotools_b200 import distributed
from oracle import oracle
from tests.golden import cases
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
distributed.enable()
fn, args, kwargs = cases.get("n3d_periodic")
_, dm = oracle.npairs_3d(*args, return_mesh=True, **kwargs)
first, last = distributed.cell1_range(dm.mesh1.ncells)
part = oracle.npairs_3d(*args, cell1_range=(first, last), **kwargs)     # stands in for this rank's GPU
total = distributed.allreduce_sum(part)
assert np.array_equal(total, oracle.npairs_3d(*args, **kwargs)), (total, part)
assert first < last and (first == 0 or last == dm.mesh1.ncells)
# several calls of one statistic: partial counts inside local_counts(), ONE all-reduce (per dtype) on exit, in place
full = oracle.npairs_3d(*args, **kwargs)
with distributed.local_counts() as partial:
    a = partial.add(np.diff(distributed.allreduce_sum(part)))           # no reduction inside the block
    b = partial.add(distributed.allreduce_sum(part.astype(np.float64)) * 0.5)
    assert np.array_equal(a, np.diff(part))
    same = partial.add(a)                                               # registered twice, reduced once
assert np.array_equal(a, np.diff(full)) and same is a and np.array_equal(b, 0.5 * full)
assert np.array_equal(distributed.allreduce_sum(part), full)            # back to one all-reduce per call
# a whole statistic over the two ranks: hb.tpcf with its pair counter served by the oracle on THIS rank's cell range
# (what the GPU engine does per rank); the counts of DD, DR and RR are all-reduced once, at the end of tpcf's block
import importlib
import halotools_b200 as hb
tp = importlib.import_module("halotools_b200.two_point_clustering.tpcf")
calls = []
def sharded_npairs_3d(a, b, rbins, period=None, num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    kw = dict(period=period, approx_cell1_size=approx_cell1_size, approx_cell2_size=approx_cell2_size)
    _, mesh = oracle.npairs_3d(a[:1], b[:1], rbins, return_mesh=True, **kw) if period is not None else oracle.npairs_3d(a, b, rbins, return_mesh=True, **kw)
    rng = distributed.cell1_range(mesh.mesh1.ncells)
    calls.append(rng)
    return distributed.allreduce_sum(oracle.npairs_3d(a, b, rbins, cell1_range=rng, **kw))
tp.npairs_3d = sharded_npairs_3d
fn, targs, tkw = cases.get("tpcf_randoms_Landy-Szalay")
xi = hb.tpcf(*targs, **tkw)
want = np.load(%(root)r + "/tests/golden/golden.npz")["tpcf_randoms_Landy-Szalay/0"]
assert np.allclose(xi, want, rtol=1e-10, atol=1e-12), (xi, want)
assert len(calls) == 3 and all(c[0] < c[1] for c in calls)
# a rank that fails inside a statistic does not leave its peer waiting in the all-reduce: both raise (ADVICE r1)
try:
    with distributed.local_counts() as partial:
        partial.add(np.zeros(3))
        if int(sys.argv[1]) == 1:
            raise ValueError("rank 1 failed")
    failed = None
except ValueError as e:
    failed = "own"
except RuntimeError as e:
    failed = "peer" if "peer rank" in str(e) else str(e)
assert failed == ("own" if int(sys.argv[1]) == 1 else "peer"), failed
dist.destroy_process_group()
print("rank", sys.argv[1], "ok", first, last)
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_host_minmax_helper_matches_numpy():
    """htb_host_minmax is pure host code of the library (no CUDA call): usable without a GPU"""
    import ctypes
    from halotools_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(5)
    a = rng.uniform(-3, 7, (300000, 3))
    lo, hi = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
    assert lib.htb_host_minmax(ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(len(a)), ctypes.c_int64(3),
                               ctypes.c_int32(3), lo, hi) == 0
    assert list(lo) == list(a.min(axis=0)) and list(hi) == list(a.max(axis=0))
    a[1234, 1] = np.nan
    assert lib.htb_host_minmax(ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(len(a)), ctypes.c_int64(3),
                               ctypes.c_int32(3), lo, hi) == 0
    assert all(np.isnan(v) for v in lo) and all(np.isnan(v) for v in hi)

    def extrema(m):
        assert lib.htb_host_minmax(ctypes.c_void_p(m.ctypes.data), ctypes.c_int64(len(m)), ctypes.c_int64(m.strides[0] // 8),
                                   ctypes.c_int32(3), lo, hi) == 0
        return list(lo), list(hi)
    # the vectorised scan of contiguous (n, 3) rows: every remainder of the 8-row trips, an infinity (its NaN detector fires
    # on inf - inf: the rows are then re-checked one by one), a NaN in the tail rows, and strided rows (the generic loop)
    for n in (1, 7, 8, 9, 1000, 100003):
        m = rng.uniform(-3, 1000, (n, 3))
        assert extrema(m) == (list(m.min(axis=0)), list(m.max(axis=0)))
    m = rng.uniform(0, 10, (100001, 3))
    m[5000, 1] = np.inf
    assert extrema(m) == (list(m.min(axis=0)), list(m.max(axis=0)))
    m[100000, 0] = np.nan
    assert all(np.isnan(v) for v in extrema(m)[0])
    m = rng.uniform(0, 10, (70000, 4))[:, :3]
    assert extrema(m) == (list(m.min(axis=0)), list(m.max(axis=0)))


def test_pbc_check_messages_on_large_samples():
    from halotools_b200.helpers import enforce_sample_respects_pbcs
    s = np.random.RandomState(6).uniform(0, 10.0, (1200000, 3))
    enforce_sample_respects_pbcs(s[:, 0], s[:, 1], s[:, 2], [10.0, 10.0, 10.0])
    s[77, 2] = -0.5
    with pytest.raises(ValueError) as err:
        enforce_sample_respects_pbcs(s[:, 0], s[:, 1], s[:, 2], [10.0, 10.0, 10.0])
    assert "negative values" in str(err.value)
    s[77, 2] = 10.5
    with pytest.raises(ValueError) as err:
        enforce_sample_respects_pbcs(s[:, 0], s[:, 1], s[:, 2], [10.0, 10.0, 10.0])
    assert "zperiod" in str(err.value)


def test_cuboid_subvolume_labels_and_jackknife_argument_errors():
    """catalog_analysis_helpers.py:330-421 semantics; tpcf_jackknife.py:601-634 / npairs_jackknife_3d.py:199-262 errors"""
    from halotools_b200.catalog_analysis_helpers import cuboid_subvolume_labels
    from halotools_b200.two_point_clustering.tpcf_jackknife import (get_subvolume_numbers, _tpcf_jackknife_process_args,
                                                                    _enclose_in_box)
    from halotools_b200.pair_counters.npairs_jackknife_3d import _process_weights_jtags
    pts = np.array([[0.0, 0.0, 0.0], [0.99, 0.99, 0.99], [1.0, 1.0, 1.0], [0.5, 0.0, 0.26], [0.24, 0.6, 0.9]])
    labels, n = cuboid_subvolume_labels(pts, [2, 2, 4], 1.0)
    assert n == 16 and list(labels) == [1, 16, 16, 10, 8]
    assert list(get_subvolume_numbers(labels, n)) == [1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 2]
    with pytest.raises(TypeError):
        cuboid_subvolume_labels(pts[:, :2], 2, 1.0)
    with pytest.raises(HalotoolsError, match="true randoms"):
        _tpcf_jackknife_process_args(S, [100], np.array([0.1, 0.2]), 2, None, None, True, True, "Natural", 1, None)
    with pytest.raises(HalotoolsError, match="Nsub"):
        _tpcf_jackknife_process_args(S, S, np.array([0.1, 0.2]), [0, 1, 1], None, 1.0, True, True, "Natural", 1, None)
    out = _tpcf_jackknife_process_args(S, [50], np.array([0.1, 0.2]), 2, None, 1.0, True, True, "Natural", 1, 43)
    assert out[4].shape == (50, 3) and np.all(out[4] <= 1.0)
    again = _tpcf_jackknife_process_args(S, [50], np.array([0.1, 0.2]), 2, None, 1.0, True, True, "Natural", 1, 43)
    assert np.array_equal(out[4], again[4])
    a, b, c, L = _enclose_in_box(S + 3.0, S + 4.0, S + 2.0)
    assert np.isclose(c.min(), 0.0) and np.all(L == L[0]) and L[0] >= b.max()
    tags = np.random.RandomState(1).randint(1, 5, len(S))
    with pytest.raises(HalotoolsError, match="jtags1 must be >= 1"):
        _process_weights_jtags(S, S, None, None, tags - 1, tags, 4)
    with pytest.raises(HalotoolsError, match="jtags2 should have same len"):
        _process_weights_jtags(S, S, None, None, tags, tags[:5], 4)
    with pytest.warns(UserWarning, match="every jackknife sample"):
        w1, w2, t1, t2 = _process_weights_jtags(S, S, None, None, np.ones(len(S), dtype=int), tags, 4)
    assert w1.dtype == np.float64 and np.all(w1 == 1.0) and t1.dtype.kind == "i"


def test_tpcf_multipole_and_s_mu_argument_errors():
    """tpcf_multipole.py:72-88 (pure host algebra) and the s_mu_tpcf validation messages (s_mu_tpcf.py:549-566)"""
    from halotools_b200.two_point_clustering.s_mu_tpcf import tpcf_multipole, _s_mu_tpcf_process_args
    mu_bins = np.linspace(0, 1, 41)
    mu_c = 0.5 * (mu_bins[:-1] + mu_bins[1:])
    xi = np.vstack([np.ones_like(mu_c), 0.5 * (3 * mu_c ** 2 - 1)])       # a pure monopole and a pure quadrupole
    assert np.allclose(tpcf_multipole(xi, mu_bins, order=0), [1.0, 0.0], atol=2e-4)
    assert np.allclose(tpcf_multipole(xi, mu_bins, order=2), [0.0, 1.0], atol=2e-3)
    with pytest.raises(ValueError, match="range"):
        _s_mu_tpcf_process_args(S, np.array([0.1, 0.2]), np.array([0.0, 1.5]), None, None, 1.0, True, True, "Natural", 1)
    with pytest.raises(ValueError, match="randoms must be provided"):
        _s_mu_tpcf_process_args(S, np.array([0.1, 0.2]), np.array([0.0, 1.0]), None, None, None, True, True, "Natural", 1)


class _FakeCudaColumn(object):
    """Stands in for a column view of a float64 CUDA tensor (no GPU here): enough surface for ``_lib.Columns``."""
    is_cuda = True

    def __init__(self, n):
        import torch
        self.shape = (n,)
        self.dtype = torch.float64

    def dim(self):
        return 1

    def stride(self, _):
        return 3

    def data_ptr(self):
        return 4096


class _FakeCudaSample(object):
    is_cuda = True

    def __init__(self, n):
        self.n = n
        self.shape = (n, 3)

    def __getitem__(self, key):
        return _FakeCudaColumn(self.n)


def test_samples_must_live_on_the_same_side_of_pcie():
    """a CUDA-tensor sample next to a numpy sample must raise before any pointer reaches the engine, and
    counters whose weights are host arrays refuse device samples."""
    dev = _FakeCudaSample(10)
    host = np.random.RandomState(0).uniform(0, 1, (10, 3))
    rb = np.array([0.01, 0.1])
    for fn, args in [(hb.npairs_3d, (rb,)), (hb.npairs_xy_z, (rb, rb)), (hb.npairs_s_mu, (rb, np.linspace(0, 1, 3))),
                     (hb.npairs_projected, (rb, 0.2))]:
        for s1, s2 in [(dev, host), (host, dev)]:
            with pytest.raises(TypeError, match="both be host arrays or both be CUDA tensors"):
                fn(s1, s2, *args, period=1.0)
    with pytest.raises(TypeError, match="takes host"):
        hb.marked_npairs_3d(dev, dev, rb, 1, period=1.0, weights1=np.ones(10), weights2=np.ones(10))
    with pytest.raises(TypeError, match="takes host"):
        hb.npairs_per_object_3d(dev, dev, rb, period=1.0)
    with pytest.raises(ValueError, match="explicit ``period``"):
        hb.npairs_3d(dev, dev, rb)
