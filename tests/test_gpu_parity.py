"""GPU parity tests (pytest -m gpu): the CUDA path, called through the public functions and hence
through the C ABI of libhalotools_b200.so, against (1) the golden vectors of the unmodified
reference, (2) the CPU oracle on the same seeded inputs, (3) size-independent properties at larger
sizes.  Integer counts must be IDENTICAL; float sums within the tolerance written in each test."""
import numpy as np
import pytest

import halotools_b200 as hb
from halotools_b200 import _lib
from oracle import oracle
from tests.gates import delta_sigma_gate, oracle_delta_sigma
from tests.golden import cases

pytestmark = pytest.mark.gpu

INT_FUNCS = ("npairs_3d", "npairs_xy_z", "npairs_s_mu", "npairs_projected", "npairs_per_object_3d")
ALL = cases.names()


def run_gpu(name, flags=0):
    fn, args, kwargs = cases.get(name)
    old = _lib.default_flags
    _lib.default_flags = flags
    try:
        return cases.flatten(getattr(hb, fn)(*args, **kwargs))
    finally:
        _lib.default_flags = old


def compare(fn, got, want, name=None):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape, (g.shape, w.shape)
        if fn in INT_FUNCS:
            assert g.dtype == np.int64
            assert np.array_equal(g, w), (g, w)
        elif fn in ("marked_npairs_3d", "marked_npairs_xy_z", "weighted_npairs_xy", "weighted_npairs_per_object_xy",
                    "total_mass_enclosed_per_cylinder", "total_mass_enclosed_in_stack_of_cylinders",
                    "surface_density_in_annulus", "surface_density_in_cylinder"):
            # float sums, order differs from the reference's serial loop: 1e-12 relative (north_star)
            assert np.allclose(g, w, rtol=1e-12, atol=0), (g, w)
        elif fn in ("npairs_jackknife_3d", "npairs_jackknife_xy_z"):
            # counts[s] = T - (A[s] + B[s]) / 2 is formed from three float sums: 1e-12 relative to the full-sample
            # counts (row 0) of the same bin; an element the reference gets as an exact 0 (both points always in the
            # removed sample) comes out at the 1e-16 level of that row
            assert np.allclose(g, w, rtol=1e-12, atol=0) or np.all(np.abs(g - w) <= 1e-12 * np.abs(w[0])[None]), \
                np.max(np.abs(g - w) / np.abs(w[0])[None])
        elif fn in ("tpcf_jackknife", "wp_jackknife", "rp_pi_tpcf_jackknife"):
            # the covariance is the scatter of the sub-sample estimators (differences of nearly equal numbers):
            # 1e-8 relative to the largest element
            assert np.allclose(g, w, rtol=1e-8, atol=1e-10 * np.max(np.abs(w)), equal_nan=True), np.max(np.abs(g - w))
        elif fn == "mean_delta_sigma":
            # cancelling difference of large sums: the SURVEY.md 8d gate 1e-12 |ref| + 1e-12 A_ik against the REFERENCE's
            # golden rows; A_ik (sum of |terms|) from the C oracle on the same inputs
            _, args, kwargs = cases.get(name)
            _, A = oracle_delta_sigma(oracle, args, kwargs)
            delta_sigma_gate(g, w, A, "golden:" + name)
        else:
            # estimators over identical counts
            # (integer counts are identical, so only the marked statistics' float sums can move the result)
            assert np.allclose(g, w, rtol=1e-12, atol=1e-12 * max(1.0, float(np.nanmax(np.abs(w)))), equal_nan=True), (g, w)


@pytest.mark.parametrize("name", ALL)
def test_matches_reference_golden(name, golden):
    fn = cases._cases()[name][0]
    compare(fn, run_gpu(name), golden(name), name)


ENGINE_INT = [n for n in ALL if cases._cases()[n][0] in INT_FUNCS and n != "n3d_c1_full"]


@pytest.mark.parametrize("name", ENGINE_INT)
@pytest.mark.parametrize("flags", [_lib.FLAG_GENERIC, _lib.FLAG_NO_CULL, _lib.FLAG_GENERIC | _lib.FLAG_NO_CULL,
                                   _lib.FLAG_NO_SYM, _lib.FLAG_NO_SYM | _lib.FLAG_NO_CULL])
def test_kernel_variants_agree(name, flags, golden):
    fn = cases._cases()[name][0]
    compare(fn, run_gpu(name, flags), golden(name), name)


def test_fast_path_taken_and_culling_reduces_work():
    fn, args, kwargs = cases.get("n3d_c1_small")
    hb.npairs_3d(*args, **kwargs)
    st = dict(_lib.last_stats)
    assert st["path"] == 1
    assert st["pairs_evaluated"] > 0
    assert st["pairs_evaluated"] <= st["pairs_reference"]
    _, dm = oracle.npairs_3d(*args, return_mesh=True, **kwargs)
    assert st["pairs_reference"] == dm.visited_pairs()
    old = _lib.default_flags
    _lib.default_flags = _lib.FLAG_NO_CULL | _lib.FLAG_NO_SYM
    try:
        hb.npairs_3d(*args, **kwargs)
        assert _lib.last_stats["pairs_evaluated"] == st["pairs_reference"]
        # symmetric auto-correlation evaluates roughly half of the pairs of the two-sided count
        _lib.default_flags = _lib.FLAG_NO_SYM
        hb.npairs_3d(*args, **kwargs)
        two_sided = _lib.last_stats["pairs_evaluated"]
    finally:
        _lib.default_flags = old
    assert st["pairs_evaluated"] < 0.7 * two_sided


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_geometries_vs_oracle(seed):
    rng = np.random.RandomState(seed)
    n1, n2 = rng.randint(1, 3000), rng.randint(1, 3000)
    L = float(rng.uniform(0.5, 300.0))
    period = [L, L * rng.uniform(1.0, 2.0), L * rng.uniform(1.0, 3.0)] if seed != 2 else None
    ext = np.array(period) if period is not None else np.array([L, L, L])
    s1 = rng.uniform(0, 1, (n1, 3)) * ext
    s2 = rng.uniform(0, 1, (n2, 3)) * ext
    rmax = L / rng.uniform(3.0, 12.0)
    rbins = np.sort(rng.uniform(0.0, rmax, rng.randint(2, 16)))
    rbins[-1] = rmax
    want = oracle.npairs_3d(s1, s2, rbins, period=period)
    got = hb.npairs_3d(s1, s2, rbins, period=period)
    assert np.array_equal(got, want)
    rp, pi = rbins, np.sort(rng.uniform(0.0, rmax, 3))
    assert np.array_equal(hb.npairs_xy_z(s1, s2, rp, pi, period=period), oracle.npairs_xy_z(s1, s2, rp, pi, period=period))


def test_dense_sample_many_tiles_vs_oracle():
    # dense enough that the refined grids (m > 1) and multi-tile columns are exercised
    s1 = cases.pts(5, 60000, 100.0)
    s2 = cases.pts(6, 80000, 100.0)
    rbins = np.logspace(-1, np.log10(10.0), 15)
    got = hb.npairs_3d(s1, s2, rbins, period=100.0)
    st = dict(_lib.last_stats)
    want = oracle.npairs_3d(s1, s2, rbins, period=100.0, num_threads=8)
    assert np.array_equal(got, want)
    assert max(st["refine2"]) > 1 and st["pairs_evaluated"] < st["pairs_reference"]


def test_empty_and_tiny_inputs():
    rbins = np.array([0.1, 0.2, 0.3])
    one = np.array([[0.5, 0.5, 0.5]])
    assert np.array_equal(hb.npairs_3d(one, one, rbins, period=1.0), [1, 1, 1])
    two = np.array([[0.05, 0.5, 0.5], [0.95, 0.5, 0.5]])
    assert np.array_equal(hb.npairs_3d(two, two, rbins, period=1.0), [2, 4, 4])
    assert np.array_equal(hb.npairs_3d(two, two, rbins, period=None), oracle.npairs_3d(two, two, rbins, period=None))
    empty = np.zeros((0, 3))
    assert np.array_equal(hb.npairs_3d(empty, one, rbins, period=1.0), [0, 0, 0])
    assert np.array_equal(hb.npairs_3d(one, empty, rbins, period=1.0), [0, 0, 0])


def test_float32_and_strided_inputs():
    s1, s2 = cases.pts(43, 500).astype(np.float32), cases.pts(44, 500)[::2]
    rbins = np.array([0.05, 0.1, 0.2])
    want = oracle.npairs_3d(s1.astype(np.float64), s2, rbins, period=1.0)
    assert np.array_equal(hb.npairs_3d(s1, s2, rbins, period=1.0), want)
    fortran = np.asfortranarray(cases.pts(45, 400))
    assert np.array_equal(hb.npairs_3d(fortran, fortran, rbins, period=1.0),
                          oracle.npairs_3d(fortran, fortran, rbins, period=1.0))


def test_cell1_ranges_are_additive():
    """The reference's parallelisation unit (cell1_tuple) is the multi-GPU shard: partial counts add up."""
    import ctypes
    from halotools_b200.pair_counters.mesh_helpers import double_mesh_geometry
    s1, s2 = cases.pts(43, 3000, 100.0), cases.pts(44, 3000, 100.0)
    rbins = np.logspace(-1, 1, 10)
    geom = double_mesh_geometry(3, [10.0] * 3, [10.0] * 3, [10.0] * 3, [100.0] * 3, True)
    g = geom.as_struct()
    c1, c2 = _lib.Columns([s1[:, 0], s1[:, 1], s1[:, 2]]), _lib.Columns([s2[:, 0], s2[:, 1], s2[:, 2]])
    total = np.zeros(len(rbins), dtype=np.int64)
    nc = geom.ncells1
    for a, b in ((0, nc // 4), (nc // 4, nc // 4), (nc // 4, nc - 7), (nc - 7, nc)):
        part = np.zeros(len(rbins), dtype=np.int64)
        _lib.run_engine("htb_npairs_3d_engine", ctypes.byref(g),
                        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
                        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
                        _lib._dp(rbins), ctypes.c_int32(len(rbins)), ctypes.c_int64(a), ctypes.c_int64(b),
                        part.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        total += part
    assert np.array_equal(total, oracle.npairs_3d(s1, s2, rbins, period=100.0))


@pytest.mark.parametrize("L,nd", [(250.0, 12), (1000.0, 33), (73.7, 129), (1.0, 7), (1000.0, 200)])
def test_mesh_cell_ids_and_offsets_match_numpy_semantics(L, nd):
    import ctypes
    from oracle.mesh import Mesh
    cs = L / nd
    edge = np.array([cs * k for k in range(nd + 1)])          # multiples of the cell size: the floor-division quirk
    rng = np.random.RandomState(3)
    near = [edge]
    for _ in range(3):                                        # ... and the 3 doubles either side of each
        near.append(np.nextafter(near[-1], 0))
    up = edge
    for _ in range(3):
        up = np.nextafter(up, 2 * L)
        near.append(up)
    x = np.concatenate(near + [rng.uniform(0, L, 5000)])
    x = np.clip(x, 0, L)
    y = rng.permutation(x)
    z = rng.permutation(x)
    m = Mesh([x, y, z], [L, L, L], [cs, cs, cs])
    ids = np.zeros(len(x), dtype=np.int64)
    cell = (ctypes.c_double * 3)(cs, cs, cs)
    ndv = (ctypes.c_int32 * 3)(nd, nd, nd)
    lib = _lib.require_gpu()
    cols = _lib.Columns([x, y, z])
    _lib.check(lib.htb_mesh_cell_ids(3, cols.ptrs[0], cols.ptrs[1], cols.ptrs[2], ctypes.c_int64(1),
                                     ctypes.c_int64(len(x)), cell, ndv,
                                     ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_uint32(0)))
    assert np.array_equal(ids, m.cell_ids)
    cii = np.zeros(m.ncells + 1, dtype=np.int64)
    _lib.check(lib.htb_mesh_cell_id_indices(3, cols.ptrs[0], cols.ptrs[1], cols.ptrs[2], ctypes.c_int64(1),
                                            ctypes.c_int64(len(x)), cell, ndv,
                                            cii.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_uint32(0)))
    assert np.array_equal(cii, m.cell_id_indices)


@pytest.mark.parametrize("switch", ["HTB_TILES_BY_THREAD=1", "HTB_FZ=8", "HTB_FZ=48", "HTB_M2=3,2,5", "HTB_M1=2,3,9",
                                    "HTB_MAXSLICES=7", "HTB_ITEMS_PER_WARP=9", "HTB_TAIL_EIGHTHS=16", "HTB_EARLY_EXIT=1",
                                    "HTB_REDO_INPLACE=1", "HTB_NO_STRADDLE=1", "HTB_NO_SYM=1",
                                    "HTB_NO_SORT_CACHE=1", "HTB_NO_STAGED_UPLOAD=1"])
def test_environment_switches_do_not_change_results(switch):
    """The A/B switches behind the sweeps under profiles/ (DESIGN: "Environment switches") select schedules, mesh
    refinements and work-item shapes - never results."""
    import os
    s1, s2 = cases.pts(21, 60000, 120.0), cases.pts(22, 90000, 120.0)
    rb = np.logspace(-1, np.log10(14.0), 13)
    w = np.linspace(0.5, 1.5, len(s1))
    want = (hb.npairs_3d(s1, s1, rb, period=120.0), hb.npairs_3d(s1, s2, rb, period=120.0),
            hb.npairs_xy_z(s1, s1, rb, [0.0, 20.0], period=120.0),
            hb.marked_npairs_3d(s1, s1, rb, 1, period=120.0, weights1=w, weights2=w))
    name, value = switch.split("=")
    os.environ[name] = value
    try:
        got = (hb.npairs_3d(s1, s1, rb, period=120.0), hb.npairs_3d(s1, s2, rb, period=120.0),
               hb.npairs_xy_z(s1, s1, rb, [0.0, 20.0], period=120.0),
               hb.marked_npairs_3d(s1, s1, rb, 1, period=120.0, weights1=w, weights2=w))
    finally:
        del os.environ[name]
    for a, b in zip(want[:3], got[:3]):
        assert np.array_equal(a, b)
    assert np.allclose(want[3], got[3], rtol=1e-12, atol=0)


def test_tile_list_by_warp_walks_long_columns_through_several_windows():
    """More than 1024 fine cells along the fast dimension (a 40 x 40 x 4000 box: 2000 reference cells per column):
    k_tiles_warp reloads its shared-memory window of column offsets on the way.  Same counts as the one-thread-per-column
    builder, as the sum over a split of sample2 and with the samples exchanged; a clustered part puts many tiles into a
    few columns."""
    import os
    period = [40.0, 40.0, 4000.0]
    rng = np.random.RandomState(31)
    uni = rng.uniform(0, 1.0, (400000, 3)) * period
    blob = np.mod(np.array([20.0, 20.0, 0.0]) + rng.normal(0, 1.0, (60000, 3)) * np.array([1.5, 1.5, 900.0]), period)
    s1 = np.concatenate([uni, blob])
    s2 = rng.uniform(0, 1.0, (300000, 3)) * period
    rb = np.logspace(-1.5, np.log10(2.0), 9)
    auto = hb.npairs_3d(s1, s1, rb, period=period)
    assert _lib.last_stats["refine1"][2] * 2000 > 1024
    cross = hb.npairs_3d(s1, s2, rb, period=period)
    os.environ["HTB_TILES_BY_THREAD"] = "1"
    try:
        assert np.array_equal(auto, hb.npairs_3d(s1, s1, rb, period=period))
        assert np.array_equal(cross, hb.npairs_3d(s1, s2, rb, period=period))
    finally:
        del os.environ["HTB_TILES_BY_THREAD"]
    assert np.array_equal(cross, hb.npairs_3d(s1, s2[:90000], rb, period=period) + hb.npairs_3d(s1, s2[90000:], rb, period=period))
    assert np.array_equal(cross, hb.npairs_3d(s2, s1, rb, period=period))
    assert auto[-1] > len(s1) and cross[-1] > 0


def test_linearity_and_symmetry_properties_large():
    """Size-independent properties at a size the oracle would take long for: counts are additive over
    a split of sample2 and symmetric under swapping the samples."""
    s1, s2 = cases.pts(11, 200000, 500.0), cases.pts(12, 300000, 500.0)
    rbins = np.logspace(-1, np.log10(25.0), 15)
    full = hb.npairs_3d(s1, s2, rbins, period=500.0)
    a = hb.npairs_3d(s1, s2[:100000], rbins, period=500.0)
    b = hb.npairs_3d(s1, s2[100000:], rbins, period=500.0)
    assert np.array_equal(a + b, full)
    assert np.array_equal(hb.npairs_3d(s2, s1, rbins, period=500.0), full)
    auto = hb.npairs_3d(s1, s1, rbins, period=500.0)
    assert auto[0] >= len(s1) and np.all(np.diff(auto) >= 0) and np.all((auto - len(s1)) % 2 == 0)


def test_marked_integer_weights_exact(golden):
    got = run_gpu("marked_grid_integer")[0]
    assert np.array_equal(got, golden("marked_grid_integer")[0])


def test_delta_sigma_rows_follow_input_order():
    gal, ptcl = cases.pts(43, 300, 1.0), cases.pts(44, 20000, 1.0)
    rpb = np.logspace(np.log10(0.02), np.log10(0.25), 8)
    base = hb.mean_delta_sigma(gal, ptcl, 1.0, rpb, period=1.0, per_object=True)
    perm = np.random.RandomState(0).permutation(len(gal))
    shuffled = hb.mean_delta_sigma(gal[perm], ptcl, 1.0, rpb, period=1.0, per_object=True)
    scale = np.max(np.abs(base))
    assert np.allclose(shuffled, base[perm], rtol=1e-10, atol=1e-12 * scale)


@pytest.mark.parametrize("name", ["ds_periodic_per_object", "ds_nonperiodic", "ds_cellsizes"])
def test_delta_sigma_general_mass_kernel_agrees(name, golden):
    """FLAG_GENERIC routes scalar masses through the per-particle-mass kernel (per-pair log)."""
    compare("mean_delta_sigma", run_gpu(name, _lib.FLAG_GENERIC), golden(name), name)


# ---------------------------------------------------------------- fast queue kernels (round 1, second half)
def _dup_points(rng, n, L, frac=0.05):
    """points with exact duplicates and shared coordinates (zero and tiny separations)"""
    s = rng.uniform(0, L, (n, 3))
    k = int(n * frac)
    s[:k] = s[k:2 * k]                     # exact duplicates: dsq == 0
    s[2 * k:3 * k, 2] = s[3 * k:4 * k, 2]  # same z: dz == 0
    s[4 * k:5 * k] = s[5 * k:6 * k] + 1e-13  # separations far below the key window
    return np.ascontiguousarray(np.mod(s, L))


@pytest.mark.parametrize("pi_bins", [[0.0, 25.0], [1e-7, 25.0], [0.0, 3.0]])
def test_fast_xy_z_path_vs_oracle_with_degenerate_separations(pi_bins):
    rng = np.random.RandomState(11)
    L = 120.0
    s1 = _dup_points(rng, 30000, L)
    s2 = np.vstack([s1[:5000], _dup_points(rng, 25000, L)])
    rp = np.logspace(-1.5, np.log10(12.0), 11)
    for a, b in ((s1, s1), (s1, s2)):
        got = hb.npairs_xy_z(a, b, rp, pi_bins, period=L)
        assert _lib.last_stats["path"] == 1, "fast (rp, pi) kernel not taken"
        want = oracle.npairs_xy_z(a, b, rp, pi_bins, period=L, num_threads=4)
        assert got.dtype == np.int64 and np.array_equal(got, want), (got - want)
        old = _lib.default_flags
        _lib.default_flags = _lib.FLAG_GENERIC
        try:
            gen = hb.npairs_xy_z(a, b, rp, pi_bins, period=L)
            assert _lib.last_stats["path"] == 0
        finally:
            _lib.default_flags = old
        assert np.array_equal(gen, want)


def test_fast_xy_z_not_taken_for_many_pi_edges():
    rng = np.random.RandomState(12)
    s = rng.uniform(0, 80.0, (5000, 3))
    rp = np.logspace(-1, 1, 8)
    pi = np.linspace(0, 20, 9)
    got = hb.npairs_xy_z(s, s, rp, pi, period=80.0)
    assert _lib.last_stats["path"] == 3, "many pi edges belong to the BinQ kernel"
    assert np.array_equal(got, oracle.npairs_xy_z(s, s, rp, pi, period=80.0, num_threads=4))


# ---------------------------------------------------------------- BinQ (binq.cu): any monotone edges, two bin axes
def _generic(call):
    old = _lib.default_flags
    _lib.default_flags = _lib.FLAG_GENERIC
    try:
        out = call()
        assert _lib.last_stats["path"] == 0
        return out
    finally:
        _lib.default_flags = old


@pytest.mark.parametrize("pi_bins", [np.linspace(0.0, 40.0, 41), [0.0, 1e-9, 1.0, 1.0 + 1e-12, 7.5, 30.0], [2.0, 5.0, 11.0]])
def test_binq_rp_pi_vs_oracle_with_degenerate_separations(pi_bins):
    """rp_pi_tpcf's counter: many pi edges, duplicate edges, zero edges, exact duplicates / shared z / 1e-13 separations"""
    rng = np.random.RandomState(21)
    L = 120.0
    s1 = _dup_points(rng, 30000, L)
    s2 = np.vstack([s1[:5000], _dup_points(rng, 25000, L)])
    rp = np.concatenate([[0.0], np.logspace(-1.5, np.log10(12.0), 13)])
    for a, b in ((s1, s1), (s1, s2)):
        got = hb.npairs_xy_z(a, b, rp, pi_bins, period=L)
        assert _lib.last_stats["path"] == 3, "BinQ kernel not taken"
        want = oracle.npairs_xy_z(a, b, rp, pi_bins, period=L, num_threads=4)
        assert got.dtype == np.int64 and got.shape == want.shape and np.array_equal(got, want), (got - want)
        assert np.array_equal(_generic(lambda: hb.npairs_xy_z(a, b, rp, pi_bins, period=L)), want)


@pytest.mark.parametrize("nmu", [2, 11, 40])
def test_binq_s_mu_vs_oracle_with_degenerate_separations(nmu):
    rng = np.random.RandomState(22)
    L = 100.0
    s1 = _dup_points(rng, 25000, L)
    s2 = np.vstack([s1[:4000], _dup_points(rng, 20000, L)])
    s_bins = np.logspace(-1, np.log10(15.0), 14)
    mu_bins = np.linspace(0.0, 1.0, nmu)
    for a, b in ((s1, s1), (s1, s2)):
        got = hb.npairs_s_mu(a, b, s_bins, mu_bins, period=L)
        assert _lib.last_stats["path"] == 3, "BinQ kernel not taken"
        want = oracle.npairs_s_mu(a, b, s_bins, mu_bins, period=L, num_threads=4)
        assert got.dtype == np.int64 and got.shape == want.shape and np.array_equal(got, want), (got - want)
        assert np.array_equal(_generic(lambda: hb.npairs_s_mu(a, b, s_bins, mu_bins, period=L)), want)


def test_binq_npairs_3d_more_than_16_bins_and_nonperiodic():
    rng = np.random.RandomState(23)
    s1 = _dup_points(rng, 20000, 90.0)
    s2 = rng.uniform(5.0, 70.0, (15000, 3))
    rbins = np.concatenate([[0.0, 1e-9], np.linspace(0.01, 9.0, 38)])
    for period in (90.0, None):
        got = hb.npairs_3d(s1, s2, rbins, period=period)
        assert _lib.last_stats["path"] == 3
        want = oracle.npairs_3d(s1, s2, rbins, period=period, num_threads=4)
        assert np.array_equal(got, want), (got - want)
    auto = hb.npairs_3d(s1, s1, rbins, period=90.0)
    assert np.array_equal(auto, oracle.npairs_3d(s1, s1, rbins, period=90.0, num_threads=4))
    old = _lib.default_flags
    _lib.default_flags = _lib.FLAG_NO_SYM | _lib.FLAG_NO_CULL
    try:
        assert np.array_equal(hb.npairs_3d(s1, s1, rbins, period=90.0), auto)
    finally:
        _lib.default_flags = old


def test_binq_dense_tiles_and_properties_large():
    """2e5 points: additivity over a split of sample2 and the pi-marginal of the (rp, pi) counts == the one-edge fast kernel"""
    rng = np.random.RandomState(24)
    L = 200.0
    s = rng.uniform(0, L, (200000, 3))
    rp = np.logspace(-1, np.log10(15.0), 12)
    pi = np.linspace(0.0, 30.0, 31)
    full = hb.npairs_xy_z(s, s, rp, pi, period=L)
    assert _lib.last_stats["path"] == 3
    a = hb.npairs_xy_z(s, s[:70000], rp, pi, period=L)
    b = hb.npairs_xy_z(s, s[70000:], rp, pi, period=L)
    assert np.array_equal(full, a + b)
    one = hb.npairs_xy_z(s, s, rp, [0.0, 30.0], period=L)
    assert _lib.last_stats["path"] == 1
    assert np.array_equal(full[:, [0, -1]], one)
    assert np.all(np.diff(full, axis=0) >= 0) and np.all(np.diff(full, axis=1) >= 0)


@pytest.mark.parametrize("wfunc", [1])
@pytest.mark.parametrize("shared", [True, False])
def test_fast_marked_path_vs_oracle(wfunc, shared):
    rng = np.random.RandomState(13)
    L = 150.0
    s1 = _dup_points(rng, 40000, L)
    w1 = rng.uniform(0.5, 1.5, len(s1))
    if shared:
        s2, w2 = s1, w1
    else:
        s2 = _dup_points(rng, 30000, L)
        w2 = rng.uniform(-1.0, 2.0, len(s2))
    rb = np.logspace(-1.2, np.log10(14.0), 13)
    got = hb.marked_npairs_3d(s1, s2, rb, wfunc, period=L, weights1=w1, weights2=w2)
    assert _lib.last_stats["path"] == 1, "fast marked kernel not taken"
    want = oracle.marked_npairs_3d(s1, s2, rb, wfunc, period=L, weights1=w1, weights2=w2, num_threads=4)
    # float sums in a different order than the reference's serial loop; mixed-sign weights: scale by the sum of |terms|
    scale = oracle.marked_npairs_3d(s1, s2, rb, wfunc, period=L, weights1=np.abs(w1), weights2=np.abs(w2), num_threads=4)
    assert np.all(np.abs(got - want) <= 1e-12 * scale), (got - want) / scale
    # integer weights: exact equality with 6 x the plain counts (reference test: test_marked_npairs_3d.py:261)
    i1, i2 = np.full(len(s1), 2.0), np.full(len(s2), 3.0)
    exact = hb.marked_npairs_3d(s1, s2, rb, wfunc, period=L, weights1=i1, weights2=i1 if shared else i2)
    n = hb.npairs_3d(s1, s2, rb, period=L)
    assert np.array_equal(exact, (4.0 if shared else 6.0) * n)


def test_fast_marked_not_taken_for_other_marks(golden):
    rng = np.random.RandomState(14)
    s = rng.uniform(0, 60.0, (3000, 3))
    w = rng.uniform(0, 1, (3000, 2))
    rb = np.logspace(-1, 1, 7)
    got = hb.marked_npairs_3d(s, s, rb, 5, period=60.0, weights1=w, weights2=w)
    assert _lib.last_stats["path"] == 3, "general marks belong to the weighted BinQ kernel"
    want = oracle.marked_npairs_3d(s, s, rb, 5, period=60.0, weights1=w, weights2=w, num_threads=4)
    assert np.allclose(got, want, rtol=1e-12, atol=0)


@pytest.mark.parametrize("kernel", ["cells", "queue"])
@pytest.mark.parametrize("nrp", [2, 3, 9, 16])
def test_fast_delta_sigma_path_vs_oracle(nrp, kernel):
    """the two uniform-mass kernels: cell-resolved (DSigmaR, path 2) and per-pair queue (DSigmaQ, path 1)"""
    import os
    rng = np.random.RandomState(15)
    L = 200.0
    gal = _dup_points(rng, 3000, L)
    ptcl = np.vstack([gal[:500], _dup_points(rng, 150000, L)])       # particles on top of galaxies: d == 0
    rp = np.logspace(-1, np.log10(20.0), nrp)
    if kernel == "queue":
        os.environ["HTB_NO_DSR"] = "1"
    try:
        got = hb.mean_delta_sigma(gal, ptcl, 2.5, rp, period=L, per_object=True)
    finally:
        os.environ.pop("HTB_NO_DSR", None)
    assert _lib.last_stats["path"] == (2 if kernel == "cells" else 1), "fast delta-sigma kernel not taken"
    want, A = oracle.mean_delta_sigma(gal, ptcl, 2.5, rp, period=L, per_object=True, num_threads=4, return_abs=True)
    delta_sigma_gate(got, want, A, "fast_path:%s:nrp%d" % (kernel, nrp))
    many = np.full(len(ptcl), 2.5)
    gen = hb.mean_delta_sigma(gal, ptcl, many, rp, period=L, per_object=True)     # per-particle masses: general kernel
    assert _lib.last_stats["path"] == 0
    delta_sigma_gate(gen, want, A, "general_mass:nrp%d" % nrp)


def test_fast3_tiles_straddling_reference_cells_keep_counts():
    """column tiles may straddle reference cells (union windows): same counts as one tile per cell"""
    import os
    rng = np.random.RandomState(16)
    L = 300.0
    s1 = rng.uniform(0, L, (60000, 3))
    s2 = rng.uniform(0, L, (90000, 3))
    rb = np.logspace(-1, np.log10(15.0), 12)
    a = hb.npairs_3d(s1, s2, rb, period=L)
    os.environ["HTB_NO_STRADDLE"] = "1"
    try:
        b = hb.npairs_3d(s1, s2, rb, period=L)
    finally:
        del os.environ["HTB_NO_STRADDLE"]
    assert np.array_equal(a, b)
    assert np.array_equal(a, oracle.npairs_3d(s1, s2, rb, period=L, num_threads=4))


def test_large_pageable_inputs_take_the_staged_upload():
    """>= 1M points from ordinary (pageable) numpy memory go through the pinned staging ring"""
    rng = np.random.RandomState(17)
    L = 400.0
    s = rng.uniform(0, L, (1500000, 3))
    rb = np.logspace(-1, np.log10(8.0), 9)
    a = hb.npairs_3d(s, s, rb, period=L)
    import os
    os.environ["HTB_NO_STAGED_UPLOAD"] = "1"
    try:
        b = hb.npairs_3d(s, s, rb, period=L)
    finally:
        del os.environ["HTB_NO_STAGED_UPLOAD"]
    assert np.array_equal(a, b)
    assert a[0] >= len(s)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_device_side_balanced_shards_add_up(world):
    """htb_set_shard: the ranks' work-balanced cell ranges partition the call (counts, sums and W_ref add up)"""
    rng = np.random.RandomState(18)
    L = 200.0
    # clustered sample1: a plain cell-count split would be badly unbalanced
    blob = np.mod(rng.normal(60.0, 6.0, (30000, 3)), L)
    s1 = np.vstack([blob, rng.uniform(0, L, (10000, 3))])
    s2 = rng.uniform(0, L, (50000, 3))
    rb = np.logspace(-1, np.log10(12.0), 11)
    w1, w2 = rng.uniform(0.5, 1.5, len(s1)), rng.uniform(0.5, 1.5, len(s2))
    full = hb.npairs_3d(s1, s2, rb, period=L)
    wfull = _lib.last_stats["pairs_reference"]
    mfull = hb.marked_npairs_3d(s1, s2, rb, 1, period=L, weights1=w1, weights2=w2)
    dfull = hb.mean_delta_sigma(s1[:5000], s2, 1.0, rb, period=L, per_object=True)
    dmean = hb.mean_delta_sigma(s1[:5000], s2, 1.0, rb, period=L)
    pfull = hb.npairs_per_object_3d(s1[:8000], s2, rb, period=L)
    tags1, tags2 = rng.randint(1, 6, len(s1)), rng.randint(1, 6, len(s2))
    jfull = hb.npairs_jackknife_3d(s1, s2, rb, tags1, tags2, 5, period=L, weights1=w1, weights2=w2)
    tot, wtot, mtot, dtot, dmtot, shares, ptot, jtot = 0, 0.0, 0.0, 0.0, 0.0, [], 0, 0.0
    try:
        for r in range(world):
            _lib.set_shard(r, world)
            ptot = ptot + hb.npairs_per_object_3d(s1[:8000], s2, rb, period=L)
            jtot = jtot + hb.npairs_jackknife_3d(s1, s2, rb, tags1, tags2, 5, period=L, weights1=w1, weights2=w2)
            tot = tot + hb.npairs_3d(s1, s2, rb, period=L)
            shares.append(_lib.last_stats["pairs_reference"])
            wtot += shares[-1]
            mtot = mtot + hb.marked_npairs_3d(s1, s2, rb, 1, period=L, weights1=w1, weights2=w2)
            dtot = dtot + hb.mean_delta_sigma(s1[:5000], s2, 1.0, rb, period=L, per_object=True)
            dmtot = dmtot + hb.mean_delta_sigma(s1[:5000], s2, 1.0, rb, period=L) * 5000.0
    finally:
        _lib.set_shard(0, 1)
    assert np.array_equal(tot, full)
    assert np.array_equal(ptot, pfull)
    assert np.all(np.abs(jtot - jfull) <= 1e-12 * np.abs(jfull[0])[None])
    assert wtot == wfull
    # balanced: no rank has more than its share plus one (heavy) cell's worth of work
    assert max(shares) <= wfull / world * 1.5, shares
    assert np.allclose(mtot, mfull, rtol=1e-12, atol=0)
    scale = np.max(np.abs(dfull))
    assert np.allclose(dtot, dfull, rtol=1e-10, atol=1e-12 * scale)
    assert np.allclose(dmtot / 5000.0, dmean, rtol=1e-10, atol=1e-12 * scale)


@pytest.mark.parametrize("maxslices", ["1", "3", "8"])
def test_tile_slices_keep_results(maxslices):
    """few tiles -> every tile is cut into column slices (independent work items): identical counts, same sums"""
    import os
    rng = np.random.RandomState(19)
    L = 150.0
    s1 = rng.uniform(0, L, (2000, 3))          # a handful of tiles only
    s2 = rng.uniform(0, L, (200000, 3))
    rb = np.logspace(-1, np.log10(14.0), 13)
    w1, w2 = rng.uniform(0.5, 1.5, len(s1)), rng.uniform(0.5, 1.5, len(s2))
    os.environ["HTB_MAXSLICES"] = maxslices
    try:
        a = hb.npairs_3d(s1, s2, rb, period=L)
        x = hb.npairs_xy_z(s1, s2, rb, [0.0, 20.0], period=L)
        sm = hb.npairs_s_mu(s1, s2, rb, np.linspace(0, 1, 6), period=L)
        m = hb.marked_npairs_3d(s1, s2, rb, 1, period=L, weights1=w1, weights2=w2)
        d = hb.mean_delta_sigma(s1, s2, 1.0, rb, period=L, per_object=True)
        dg = hb.mean_delta_sigma(s1[:300], s2, np.full(len(s2), 1.0), rb, period=L, per_object=True)
        auto = hb.npairs_3d(s1, s1, rb, period=L)
    finally:
        del os.environ["HTB_MAXSLICES"]
    assert np.array_equal(a, oracle.npairs_3d(s1, s2, rb, period=L, num_threads=4))
    assert np.array_equal(x, oracle.npairs_xy_z(s1, s2, rb, [0.0, 20.0], period=L, num_threads=4))
    assert np.array_equal(sm, oracle.npairs_s_mu(s1, s2, rb, np.linspace(0, 1, 6), period=L, num_threads=4))
    assert np.array_equal(auto, oracle.npairs_3d(s1, s1, rb, period=L, num_threads=4))
    assert np.allclose(m, oracle.marked_npairs_3d(s1, s2, rb, 1, period=L, weights1=w1, weights2=w2, num_threads=4),
                       rtol=1e-12, atol=0)
    want, A = oracle.mean_delta_sigma(s1, s2, 1.0, rb, period=L, per_object=True, num_threads=4, return_abs=True)
    delta_sigma_gate(d, want, A, "slices:" + maxslices)
    delta_sigma_gate(dg, want[:300], A[:300], "slices_general:" + maxslices)


def test_delta_sigma_device_resident_inputs_and_column_sums():
    """CUDA tensors in (no host copy of the samples); per_object=False returns the device-side column sums / N"""
    import torch
    rng = np.random.RandomState(20)
    L = 120.0
    gal, ptcl = rng.uniform(0, L, (4000, 3)), rng.uniform(0, L, (120000, 3))
    rp = np.logspace(-1, 1, 9)
    want, A = oracle.mean_delta_sigma(gal, ptcl, 1.5, rp, period=L, per_object=True, num_threads=4, return_abs=True)
    gd, pd = torch.from_numpy(gal).cuda(), torch.from_numpy(ptcl).cuda()
    rows = hb.mean_delta_sigma(gd, pd, 1.5, rp, period=L, per_object=True)
    delta_sigma_gate(rows, want, A, "device_rows")
    mean_dev = hb.mean_delta_sigma(gd, pd, 1.5, rp, period=L)
    mean_host = hb.mean_delta_sigma(gal, ptcl, 1.5, rp, period=L)
    assert mean_dev.shape == (len(rp) - 1,)
    delta_sigma_gate(mean_dev, np.mean(want, axis=0), np.mean(A, axis=0), "device_mean")
    delta_sigma_gate(mean_host, np.mean(want, axis=0), np.mean(A, axis=0), "host_mean")
    many = hb.mean_delta_sigma(gal, ptcl, np.full(len(ptcl), 1.5), rp, period=L)     # general-mass kernel, column sums
    delta_sigma_gate(many, np.mean(want, axis=0), np.mean(A, axis=0), "general_mass_mean")


def test_upload_cache_reuses_only_caller_owned_arrays():
    """inside upload_cache() a sample is uploaded once; temporaries (float32 conversions, non-periodic shifts)
    never take part, so recycled host addresses cannot alias stale device copies"""
    rng = np.random.RandomState(21)
    L = 90.0
    a, b = rng.uniform(0, L, (40000, 3)), rng.uniform(0, L, (50000, 3))
    rb = np.logspace(-1, 1, 8)
    want_ab = oracle.npairs_3d(a, b, rb, period=L, num_threads=4)
    want_bb = oracle.npairs_3d(b, b, rb, period=L, num_threads=4)
    with _lib.upload_cache():
        assert np.array_equal(hb.npairs_3d(a, b, rb, period=L), want_ab)
        first = _lib.last_stats["ms_h2d"]
        assert np.array_equal(hb.npairs_3d(b, b, rb, period=L), want_bb)      # b is already there
        assert np.array_equal(hb.npairs_3d(a, b, rb, period=L), want_ab)
        # temporaries: float32 input (converted copy) and the non-periodic path (shifted copies), several times
        for seed in range(3):
            c = np.random.RandomState(seed).uniform(0, L, (40000, 3))
            assert np.array_equal(hb.npairs_3d(c.astype(np.float32), b, rb, period=L),
                                  oracle.npairs_3d(c.astype(np.float32).astype(np.float64), b, rb, period=L, num_threads=4))
            assert np.array_equal(hb.npairs_3d(c, b, rb), oracle.npairs_3d(c, b, rb, num_threads=4))
    assert first >= 0.0
    # the estimator drivers use it
    xi = hb.tpcf(a, rb, randoms=b, period=L, estimator="Landy-Szalay")
    DD, DR, RR = (np.diff(oracle.npairs_3d(x, y, rb, period=L, num_threads=4)) for x, y in ((a, a), (a, b), (b, b)))
    from halotools_b200.two_point_clustering.tpcf_estimators import _TP_estimator
    assert np.allclose(xi, _TP_estimator(DD, DR, RR, len(a), len(a), len(b), len(b), "Landy-Szalay"), rtol=1e-12)


@pytest.mark.parametrize("case", ["dense", "sparse", "wide_bins", "nonperiodic", "rect"])
def test_cell_resolved_delta_sigma_vs_oracle(case):
    """DSigmaR decides the annulus per (galaxy, particle cell): dense / sparse cells, two wide bins (cells that hold a
    galaxy cross one edge), non-periodic enclosing box, non-square box"""
    rng = np.random.RandomState(22)
    L, period = 300.0, 300.0
    ngal, nptcl, rp = 4000, 600000, np.logspace(-1, np.log10(25.0), 13)
    if case == "sparse":
        nptcl = 20000
    elif case == "wide_bins":
        rp = np.array([2.0, 30.0])
    elif case == "rect":
        period = [300.0, 200.0, 100.0]
    gal = rng.uniform(0, 1, (ngal, 3)) * (period if case == "rect" else L)
    ptcl = rng.uniform(0, 1, (nptcl, 3)) * (period if case == "rect" else L)
    # a clump of particles right on top of / very close to some galaxies
    ptcl[:200] = gal[:200]
    ptcl[200:400, :2] = gal[200:400, :2] + 1e-9
    kw = {} if case == "nonperiodic" else {"period": period}
    g1, p1 = (gal.copy(), ptcl.copy()) if case == "nonperiodic" else (gal, ptcl)      # the non-periodic path shifts in place
    got = hb.mean_delta_sigma(g1, p1, 0.7, rp, per_object=True, **kw)
    assert _lib.last_stats["path"] == 2, "cell-resolved kernel not taken"
    g2, p2 = (gal.copy(), ptcl.copy()) if case == "nonperiodic" else (gal, ptcl)
    want, A = oracle.mean_delta_sigma(g2, p2, 0.7, rp, per_object=True, num_threads=8, return_abs=True, **kw)
    delta_sigma_gate(got, want, A, "cell_resolved:" + case)


# ---------------------------------------------------------------- SURVEY 8(f) rank 2 counters (BinQ modes 1 and 2)
def test_npairs_projected_vs_oracle_and_xy_z_column():
    rng = np.random.RandomState(31)
    L = 150.0
    s1 = _dup_points(rng, 40000, L)
    s2 = np.vstack([s1[:6000], _dup_points(rng, 30000, L)])
    rp = np.logspace(-1.5, np.log10(14.0), 12)
    for a, b in ((s1, s1), (s1, s2)):
        got = hb.npairs_projected(a, b, rp, 35.0, period=L)
        assert _lib.last_stats["path"] == 1, "one pi edge: the fast (rp, pi) kernel"
        want = oracle.npairs_projected(a, b, rp, 35.0, period=L)
        assert got.dtype == np.int64 and np.array_equal(got, want), (got - want)
    with pytest.raises(ValueError, match="pi_max"):
        hb.npairs_projected(s1, s1, rp, 60.0, period=L)


def test_npairs_per_object_3d_vs_oracle_rows_in_input_order():
    rng = np.random.RandomState(32)
    L = 100.0
    s1 = _dup_points(rng, 30000, L)
    s2 = np.vstack([s1[:4000], _dup_points(rng, 26000, L)])
    rbins = np.concatenate([[0.0], np.logspace(-2, np.log10(9.0), 17)])
    for a, b in ((s1, s1), (s1, s2)):
        got = hb.npairs_per_object_3d(a, b, rbins, period=L)
        want = oracle.npairs_per_object_3d(a, b, rbins, period=L)
        assert got.dtype == np.int64 and got.shape == (len(a), len(rbins)) and np.array_equal(got, want)
        # column sums are npairs_3d
        assert np.array_equal(got.sum(axis=0), hb.npairs_3d(a, b, rbins, period=L))
    sub = rng.permutation(len(s1))[:500]
    assert np.array_equal(hb.npairs_per_object_3d(s1[sub], s2, rbins, period=L),
                          hb.npairs_per_object_3d(s1, s2, rbins, period=L)[sub])
    nonper = hb.npairs_per_object_3d(s1[:3000], s2[:5000], rbins, period=None)
    assert np.array_equal(nonper, oracle.npairs_per_object_3d(s1[:3000], s2[:5000], rbins, period=None))


@pytest.mark.parametrize("wid", [1, 5, 9, 13])
def test_marked_npairs_xy_z_vs_oracle(wid):
    rng = np.random.RandomState(33)
    L = 120.0
    s1 = _dup_points(rng, 25000, L)
    s2 = np.vstack([s1[:3000], _dup_points(rng, 20000, L)])
    w1, w2 = cases.weights(1, len(s1), wid), cases.weights(2, len(s2), wid)
    rp = np.logspace(-1, np.log10(10.0), 9)
    pi = np.linspace(0.0, 24.0, 13)
    for a, b, wa, wb in ((s1, s1, w1, w1), (s1, s2, w1, w2)):
        got = hb.marked_npairs_xy_z(a, b, rp, pi, period=L, weights1=wa, weights2=wb, weight_func_id=wid)
        want = oracle.marked_npairs_xy_z(a, b, rp, pi, period=L, weights1=wa, weights2=wb, weight_func_id=wid)
        assert got.shape == want.shape and np.allclose(got, want, rtol=1e-12, atol=0), np.max(np.abs(got / want - 1))
    # unit weights, product marks: the sums are the integer counts
    ones = np.ones(len(s1))
    unit = hb.marked_npairs_xy_z(s1, s1, rp, pi, period=L, weights1=ones, weights2=ones, weight_func_id=1)
    assert np.array_equal(unit, hb.npairs_xy_z(s1, s1, rp, pi, period=L).astype(float))
    with pytest.raises(hb.HalotoolsError):
        hb.marked_npairs_xy_z(s1, s1, rp, pi, period=L)            # the reference's default id 0 is not recognised


def test_marked_npairs_3d_general_marks_take_binq_and_match_oracle():
    rng = np.random.RandomState(34)
    L = 100.0
    s1 = _dup_points(rng, 25000, L)
    s2 = _dup_points(rng, 20000, L)
    rb = np.logspace(-1, 1, 12)
    for wid in (2, 6, 12, 16):
        w1, w2 = cases.weights(3, len(s1), wid), cases.weights(4, len(s2), wid)
        got = hb.marked_npairs_3d(s1, s2, rb, wid, period=L, weights1=w1, weights2=w2)
        assert _lib.last_stats["path"] == 3
        want = oracle.marked_npairs_3d(s1, s2, rb, wid, period=L, weights1=w1, weights2=w2, num_threads=4)
        assert np.allclose(got, want, rtol=1e-12, atol=0), (wid, np.max(np.abs(got / want - 1)))
        gen = _generic(lambda: hb.marked_npairs_3d(s1, s2, rb, wid, period=L, weights1=w1, weights2=w2))
        assert np.allclose(gen, want, rtol=1e-12, atol=0)


def test_weighted_npairs_xy_vs_oracle():
    rng = np.random.RandomState(35)
    L = 300.0
    g = rng.uniform(0, L, (20000, 2))
    p = np.vstack([g[:2000], rng.uniform(0, L, (150000, 2))])
    m = rng.uniform(0.0, 2.0, len(p))
    rp = np.logspace(-1, np.log10(25.0), 14)
    # float masses at a size where the REFERENCE's serial running sum (sqrt(n) eps for n terms) stays below 1e-12;
    # at the full size only integer masses (every partial sum exact in f64) are compared, and exactly
    got = hb.weighted_npairs_xy(g[:5000], p[:40000], m[:40000], rp, period=L)
    assert _lib.last_stats["path"] == 3
    want = oracle.weighted_npairs_xy(g[:5000], p[:40000], m[:40000], rp, period=L)
    assert np.allclose(got, want, rtol=1e-12, atol=0), np.max(np.abs(got / want - 1))
    ints = rng.randint(1, 5, len(p)).astype(float)                  # integer masses: exact
    assert np.array_equal(hb.weighted_npairs_xy(g, p, ints, rp, period=L), oracle.weighted_npairs_xy(g, p, ints, rp, period=L))
    nonper = hb.weighted_npairs_xy(g[:3000], p[:30000], m[:30000], rp, period=None)
    assert np.allclose(nonper, oracle.weighted_npairs_xy(g[:3000], p[:30000], m[:30000], rp, period=None), rtol=1e-12)


def test_device_minmax_matches_numpy():
    import ctypes
    import torch
    rng = np.random.RandomState(36)
    for n in (1, 31, 1000, 300001):
        a = rng.uniform(-5.0, 7.0, (n, 3))
        t = torch.from_numpy(a).cuda()
        lo, hi = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
        _lib.check(_lib.load().htb_device_minmax(ctypes.c_void_p(int(t.data_ptr())), ctypes.c_int64(n), ctypes.c_int64(3),
                                                 ctypes.c_int32(3), lo, hi))
        assert np.array_equal(np.array(lo[:]), a.min(axis=0)) and np.array_equal(np.array(hi[:]), a.max(axis=0))
        # two of the three columns (the 2-D engines), NaN anywhere poisons every extremum
        _lib.check(_lib.load().htb_device_minmax(ctypes.c_void_p(int(t.data_ptr())), ctypes.c_int64(n), ctypes.c_int64(3),
                                                 ctypes.c_int32(2), lo, hi))
        assert np.array_equal(np.array(lo[:2]), a.min(axis=0)[:2]) and np.array_equal(np.array(hi[:2]), a.max(axis=0)[:2])
        a[n // 2, 1] = np.nan
        t = torch.from_numpy(a).cuda()
        _lib.check(_lib.load().htb_device_minmax(ctypes.c_void_p(int(t.data_ptr())), ctypes.c_int64(n), ctypes.c_int64(3),
                                                 ctypes.c_int32(3), lo, hi))
        assert all(np.isnan(v) for v in lo[:]) and all(np.isnan(v) for v in hi[:])
    # the front-end check on device tensors raises the reference's messages
    s = torch.from_numpy(rng.uniform(0, 10.0, (5000, 3))).cuda()
    s[17, 2] = 11.0
    with pytest.raises(ValueError, match="zperiod"):
        hb.mean_delta_sigma(s, s, 1.0, np.logspace(-1, 0, 4), period=10.0)


# ---------------------------------------------------------------- jackknife pair counters
@pytest.mark.parametrize("auto", [True, False])
def test_npairs_jackknife_3d_vs_oracle(auto):
    rng = np.random.RandomState(41)
    L, ns = 100.0, 27
    s1 = _dup_points(rng, 12000, L)
    s2 = s1 if auto else np.vstack([s1[:2000], _dup_points(rng, 9000, L)])
    w1 = rng.uniform(0.5, 1.5, len(s1))
    w2 = w1 if auto else rng.uniform(0.5, 1.5, len(s2))
    # spatial tags (3 x 3 x 3 sub-volumes), as tpcf_jackknife makes them
    tag = lambda s: 1 + (np.minimum((s // (L / 3)).astype(int), 2) * np.array([9, 3, 1])).sum(axis=1)
    t1, t2 = tag(s1), tag(s2)
    rbins = np.logspace(-1, np.log10(9.0), 12)
    got = hb.npairs_jackknife_3d(s1, s2, rbins, t1, t2, ns, period=L, weights1=w1, weights2=w2)
    want = oracle.npairs_jackknife_3d(s1, s2, rbins, t1, t2, ns, period=L, weights1=w1, weights2=w2)
    assert got.shape == want.shape == (ns + 1, len(rbins))
    assert np.all(np.abs(got - want) <= 1e-12 * np.abs(want[0])[None]), np.max(np.abs(got - want) / want[0][None])
    # row 0 is the plain weighted count
    full = hb.marked_npairs_3d(s1, s2, rbins, 1, period=L, weights1=w1, weights2=w2)
    assert np.allclose(got[0], full, rtol=1e-12)
    # unit weights: every entry is a multiple of 1/2 and exact
    ones1, ones2 = np.ones(len(s1)), np.ones(len(s2))
    gi = hb.npairs_jackknife_3d(s1, s2, rbins, t1, t2, ns, period=L, weights1=ones1, weights2=ones2)
    wi = oracle.npairs_jackknife_3d(s1, s2, rbins, t1, t2, ns, period=L, weights1=ones1, weights2=ones2)
    assert np.array_equal(gi, wi)
    assert np.array_equal(gi[0], hb.npairs_3d(s1, s2, rbins, period=L).astype(float))


def test_npairs_jackknife_xy_z_vs_oracle_and_errors():
    rng = np.random.RandomState(42)
    L, ns = 120.0, 8
    s1 = _dup_points(rng, 10000, L)
    s2 = _dup_points(rng, 8000, L)
    t1, t2 = rng.randint(1, ns + 1, len(s1)), rng.randint(1, ns + 1, len(s2))
    rp = np.logspace(-1, np.log10(10.0), 9)
    pi = np.array([0.0, 5.0, 25.0])
    for per in (L, None):
        got = hb.npairs_jackknife_xy_z(s1, s2, rp, pi, t1, t2, ns, period=per)
        want = oracle.npairs_jackknife_xy_z(s1, s2, rp, pi, t1, t2, ns, period=per)
        assert got.shape == want.shape == (ns + 1, len(rp), len(pi))
        assert np.array_equal(got, want)                       # unit weights: exact halves
    with pytest.raises(hb.HalotoolsError, match="jtags1 must be <= N_samples"):
        hb.npairs_jackknife_xy_z(s1, s2, rp, pi, t1 + 1, t2, ns, period=L)
    with pytest.raises(hb.HalotoolsError, match="weights2 should have same len"):
        hb.npairs_jackknife_3d(s1, s2, rp, t1, t2, ns, period=L, weights2=np.ones(5))


def test_weighted_npairs_per_object_xy_vs_oracle():
    rng = np.random.RandomState(37)
    L = 300.0
    g = rng.uniform(0, L, (6000, 2))
    p = np.vstack([g[:1000], rng.uniform(0, L, (60000, 2))])
    m = rng.uniform(0.0, 2.0, len(p))
    rp = np.logspace(-1, np.log10(25.0), 14)
    got = hb.weighted_npairs_per_object_xy(g, p, m, rp, period=L)
    assert _lib.last_stats["path"] == 3
    want = oracle.weighted_npairs_per_object_xy(g, p, m, rp, period=L)
    assert got.shape == want.shape == (len(g), len(rp))
    # each row sums <= a few thousand terms: 1e-12 relative, exact zeros stay zero
    assert np.allclose(got, want, rtol=1e-12, atol=0), np.nanmax(np.abs(got / want - 1))
    assert np.allclose(got.sum(axis=0), hb.weighted_npairs_xy(g, p, m, rp, period=L), rtol=1e-11)
    ints = rng.randint(1, 5, len(p)).astype(float)
    assert np.array_equal(hb.weighted_npairs_per_object_xy(g, p, ints, rp, period=L),
                          oracle.weighted_npairs_per_object_xy(g, p, ints, rp, period=L))
    sub = rng.permutation(len(g))[:300]
    assert np.array_equal(hb.weighted_npairs_per_object_xy(g[sub], p, ints, rp, period=L),
                          hb.weighted_npairs_per_object_xy(g, p, ints, rp, period=L)[sub])


def test_large_per_object_outputs_take_the_staged_download():
    """per-object tables >= 2M values go to pageable memory through the pinned ring: same rows as the direct copy"""
    import os
    rng = np.random.RandomState(38)
    L = 200.0
    s = rng.uniform(0, L, (150000, 3))
    rbins = np.logspace(-1, 1, 15)
    staged = hb.npairs_per_object_3d(s, s, rbins, period=L)
    assert staged.shape == (150000, 15) and staged.size >= 8 * 256 * 1024
    os.environ["HTB_NO_STAGED_UPLOAD"] = "1"
    try:
        direct = hb.npairs_per_object_3d(s, s, rbins, period=L)
    finally:
        del os.environ["HTB_NO_STAGED_UPLOAD"]
    assert np.array_equal(staged, direct)
    assert np.array_equal(staged.sum(axis=0), hb.npairs_3d(s, s, rbins, period=L))
    sub = rng.permutation(len(s))[:2000]
    assert np.array_equal(staged[sub], oracle.npairs_per_object_3d(s[sub], s, rbins, period=L))
    ptcl = rng.uniform(0, L, (400000, 3))
    rp = np.logspace(-1, 1, 15)
    rows = hb.mean_delta_sigma(s, ptcl, 1.0, rp, period=L, per_object=True)
    assert rows.shape == (150000, 14)
    assert np.allclose(rows.mean(axis=0), hb.mean_delta_sigma(s, ptcl, 1.0, rp, period=L), rtol=1e-9, atol=1e-12 * np.max(np.abs(rows)))


def test_odd_bin_arrays_keep_the_reference_scan_semantics():
    """len-2 bins need not increase (npairs_3d.py:175-182): the reference's top-down scan stops at the first failing
    edge, i.e. it counts against the suffix minima; very many bins leave BinQ for the literal-scan kernel"""
    rng = np.random.RandomState(39)
    L = 60.0
    s1, s2 = _dup_points(rng, 6000, L), _dup_points(rng, 5000, L)
    for rbins in ([3.0, 1.0], [2.0, 2.0], [0.0, 2.5]):
        got = hb.npairs_3d(s1, s2, rbins, period=L)
        assert np.array_equal(got, oracle.npairs_3d(s1, s2, rbins, period=L)), rbins
    got = hb.npairs_xy_z(s1, s2, [2.0, 1.0], [5.0, 0.5], period=L)
    assert np.array_equal(got, oracle.npairs_xy_z(s1, s2, [2.0, 1.0], [5.0, 0.5], period=L))
    many = np.linspace(0.05, 5.0, 200)
    got = hb.npairs_3d(s1, s2, many, period=L)
    assert _lib.last_stats["path"] == 0, "more than 127 edges: the literal-scan kernel"
    assert np.array_equal(got, oracle.npairs_3d(s1, s2, many, period=L, num_threads=4))
    wide = np.linspace(0.05, 5.0, 120)
    got = hb.npairs_3d(s1, s2, wide, period=L)
    assert _lib.last_stats["path"] == 3
    assert np.array_equal(got, oracle.npairs_3d(s1, s2, wide, period=L, num_threads=4))


# ------------------------------------------------------------------ K3 on the device (statistics: one synchronisation per call)
def _host_statistic(counter_name, module, fn, *args, **kwargs):
    """the same statistic through its host restatement: the pair counter stripped of its ``enqueue`` entry"""
    import importlib
    mod = importlib.import_module("halotools_b200.two_point_clustering." + module)
    real = getattr(mod, counter_name)

    def plain(*a, **k):
        return real(*a, **k)
    setattr(mod, counter_name, plain)
    try:
        return getattr(hb, fn)(*args, **kwargs)
    finally:
        setattr(mod, counter_name, real)


@pytest.mark.parametrize("estimator", ["Natural", "Davis-Peebles", "Hewett", "Hamilton", "Landy-Szalay"])
@pytest.mark.parametrize("cross", [False, True])
def test_device_estimator_is_the_numpy_formula_bit_for_bit(estimator, cross):
    # tpcf_estimators.py:14-119 evaluated by htb_tp_estimator on the device tables vs numpy on the host counts
    rng = np.random.RandomState(5)
    s1, s2, ran = rng.uniform(0, 60.0, (3000, 3)), rng.uniform(0, 60.0, (2500, 3)), rng.uniform(0, 60.0, (9000, 3))
    rbins = np.logspace(-0.3, 1.0, 9)
    kw = dict(randoms=ran, period=60.0, estimator=estimator)
    if cross:
        if estimator in ("Davis-Peebles", "Hewett"):
            pytest.skip("no cross-correlation form in the reference")
        kw["sample2"] = s2
    got = hb.tpcf(s1, rbins, **kw)
    want = _host_statistic("npairs_3d", "tpcf", "tpcf", s1, rbins, **kw)
    got = got if isinstance(got, tuple) else (got,)
    want = want if isinstance(want, tuple) else (want,)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w, equal_nan=True), (g, w)


def test_device_statistics_analytic_randoms_wp_and_zero_division():
    rng = np.random.RandomState(6)
    s1, s2 = rng.uniform(0, 80.0, (4000, 3)), rng.uniform(0, 80.0, (3000, 3))
    rp = np.logspace(-0.5, 1.0, 8)
    # analytic randoms (float differential tables on the device), auto + cross
    for fn, module, counter, args in (("tpcf", "tpcf", "npairs_3d", (s1, rp)),
                                      ("rp_pi_tpcf", "rp_pi_tpcf", "npairs_xy_z", (s1, rp, np.linspace(0, 20, 6))),
                                      ("wp", "rp_pi_tpcf", "npairs_xy_z", (s1, rp, 20.0))):
        kw = dict(sample2=s2, period=80.0)
        got = getattr(hb, fn)(*args, **kw)
        want = _host_statistic(counter, module, fn, *args, **kw)
        assert len(got) == 3
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w, equal_nan=True), (fn, g, w)
    # an empty RR bin raises the reference's ValueError (tpcf_estimators.py:165-183)
    ran = rng.uniform(0, 80.0, (50, 3))
    with pytest.raises(ValueError, match="zero RR pairs"):
        hb.tpcf(s1, np.array([1e-4, 2e-4, 1.0]), randoms=ran, period=80.0, estimator="Landy-Szalay")
    # device-resident samples: same numbers as host arrays
    import torch
    d1, dr = torch.from_numpy(s1).cuda(), torch.from_numpy(rng.uniform(0, 80.0, (20000, 3))).cuda()
    a = hb.tpcf(d1, rp, randoms=dr, period=80.0, estimator="Landy-Szalay")
    b = hb.tpcf(s1, rp, randoms=dr.cpu().numpy(), period=80.0, estimator="Landy-Szalay")
    assert np.array_equal(a, b)


def test_asynchronous_engine_calls_leave_counts_on_the_device():
    # HTB_FLAG_DEVICE_OUTPUT through the front-ends' enqueue entries: every kernel path, no host synchronisation until the read
    import torch
    rng = np.random.RandomState(7)
    s1, s2 = rng.uniform(0, 50.0, (5000, 3)), rng.uniform(0, 50.0, (7000, 3))
    rb = np.logspace(-1, 1, 10)
    many = np.logspace(-1, 1, 24)                      # > 16 bins: BinQ
    pi = np.linspace(0, 12.0, 7)
    _lib.async_count_times()                           # (empty the event ring)
    with torch.cuda.stream(_lib.engine_stream()):
        t1 = torch.zeros(len(rb), dtype=torch.int64, device="cuda")
        t2 = torch.zeros(len(many), dtype=torch.int64, device="cuda")
        t3 = torch.zeros(len(rb) * 2, dtype=torch.int64, device="cuda")
        t4 = torch.zeros(len(rb) * len(pi), dtype=torch.int64, device="cuda")
        keep = [hb.npairs_3d.enqueue(t1, s1, s2, rb, period=50.0), hb.npairs_3d.enqueue(t2, s1, s1, many, period=50.0),
                hb.npairs_xy_z.enqueue(t3, s1, s2, rb, [0.0, 9.0], period=50.0),
                hb.npairs_xy_z.enqueue(t4, s1, s2, rb, pi, period=50.0)]
        old = _lib.default_flags
        _lib.default_flags = _lib.FLAG_GENERIC
        try:
            t5 = torch.zeros(len(rb), dtype=torch.int64, device="cuda")
            keep.append(hb.npairs_3d.enqueue(t5, s1, s2, rb, period=50.0))
        finally:
            _lib.default_flags = old
        got = [t.cpu().numpy() for t in (t1, t2, t3, t4, t5)]
    spans = _lib.async_kernel_spans()                  # the kernels' own device-side stamps (ask before the ring is reset)
    times = _lib.async_count_times()
    assert len(times) == 5 and all(t > 0 for t in times)
    # first warp in -> last warp out lies inside the CUDA-event bracket of the launch (the stamps tick in ~1 us steps)
    assert len(spans) == 5 and all(0 < s <= t + 0.01 for s, t in zip(spans, times)), (spans, times)
    assert np.array_equal(got[0], hb.npairs_3d(s1, s2, rb, period=50.0))
    assert np.array_equal(got[1], hb.npairs_3d(s1, s1, many, period=50.0))
    assert np.array_equal(got[2].reshape(len(rb), 2), hb.npairs_xy_z(s1, s2, rb, [0.0, 9.0], period=50.0))
    assert np.array_equal(got[3].reshape(len(rb), len(pi)), hb.npairs_xy_z(s1, s2, rb, pi, period=50.0))
    assert np.array_equal(got[4], got[0])


# ------------------------------------------------------------------ 8f-4: the input step on the device
class _FlatLCDM(object):
    """E(z) of a flat LCDM cosmology (what astropy's efunc returns up to radiation)."""
    def __init__(self, om=0.3):
        self.om = om

    def efunc(self, z):
        return np.sqrt(self.om * (1.0 + z) ** 3 + 1.0 - self.om)


def test_input_step_on_the_device_is_the_host_function_bit_for_bit():
    # catalog_analysis_helpers.py:108-327 - device tensors in, device (Npts, 3) sample out, no coordinate crosses PCIe
    import torch
    rng = np.random.RandomState(11)
    n, L = 200000, 250.0
    x, y, z = (rng.uniform(-0.3 * L, 1.3 * L, n) for _ in range(3))       # outside the box on both sides: np.mod wraps
    x[:5] = [0.0, L, -L, 2 * L, -0.0]
    v = rng.normal(0.0, 400.0, n)
    dx, dy, dz, dv = (torch.from_numpy(a).cuda() for a in (x, y, z, v))
    cosmo = _FlatLCDM()
    for kw in (dict(period=L), dict(period=[L, 2 * L, 3 * L]), dict(),
               dict(period=L, velocity=v, velocity_distortion_dimension="z"),
               dict(period=L, velocity=v, velocity_distortion_dimension="x", redshift=1.5, cosmology=cosmo),
               dict(velocity=v, velocity_distortion_dimension="y", redshift=0.7, cosmology=cosmo)):
        want = hb.return_xyz_formatted_array(x, y, z, **kw)
        dkw = dict(kw)
        if "velocity" in dkw:
            dkw["velocity"] = dv
        got = hb.return_xyz_formatted_array(dx, dy, dz, **dkw)
        assert got.is_cuda and tuple(got.shape) == (n, 3)
        assert np.array_equal(got.cpu().numpy(), want, equal_nan=True), kw
    mask = x > 10.0
    got = hb.return_xyz_formatted_array(dx, dy, dz, period=L, mask=torch.from_numpy(mask).cuda())
    assert np.array_equal(got.cpu().numpy(), hb.return_xyz_formatted_array(x, y, z, period=L, mask=mask))
    for Lbox in (None, L):
        want = hb.apply_zspace_distortion(np.mod(z, L), v, 0.8, cosmo, Lbox=Lbox)
        got = hb.apply_zspace_distortion(torch.from_numpy(np.mod(z, L)).cuda(), dv, 0.8, cosmo, Lbox=Lbox)
        assert np.array_equal(got.cpu().numpy(), want)
    # the device sample feeds the statistics directly
    pos_d = hb.return_xyz_formatted_array(dx, dy, dz, period=L, velocity=dv, velocity_distortion_dimension="z")
    pos_h = hb.return_xyz_formatted_array(x, y, z, period=L, velocity=v, velocity_distortion_dimension="z")
    rp = np.logspace(-0.5, 1.2, 9)
    a = hb.wp(pos_d, rp, 40.0, period=L)
    b = hb.wp(pos_h, rp, 40.0, period=L)
    assert np.array_equal(a, b)


def test_fast_path_keeps_counts_for_separations_within_a_few_ulp_of_round_and_generic_edges():
    # separations within a few ulp of an edge - round edges (squares with zero low bits: the edge sits at the bottom of
    # its 32-bit key cell) and generic ones - must be decided exactly as the reference decides them
    rng = np.random.RandomState(21)
    n = 40000
    s = rng.uniform(0, 60.0, (n, 3))
    # pairs at (nearly) exactly the edge separations: partner = point + edge * unit vector, nudged by a few ulp
    edges = np.array([0.5, 1.0, 2.0, 3.0, 5.0])
    base = s[:2000].copy()
    u = rng.normal(size=(2000, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    part = np.mod(base + u * edges[rng.randint(0, 5, 2000)][:, None] * (1.0 + rng.randint(-4, 5, 2000)[:, None] * 2.2e-16), 60.0)
    s2 = np.vstack([s, part])
    for rb in (edges, edges * 1.0000001234, np.logspace(-0.5, 0.8, 12)):
        want = oracle.npairs_3d(s2, s2, rb, period=60.0, num_threads=8)
        got = hb.npairs_3d(s2, s2, rb, period=60.0)
        assert _lib.last_stats["path"] == 1
        assert np.array_equal(got, want), (rb, got, want)
        wx = oracle.npairs_xy_z(s2, s, rb, [0.0, 7.0], period=60.0, num_threads=8)
        gx = hb.npairs_xy_z(s2, s, rb, [0.0, 7.0], period=60.0)
        assert np.array_equal(gx, wx)
    w = rng.randint(1, 4, len(s2)).astype(float)
    for rb in (edges, edges * 1.0000001234):
        wm = oracle.marked_npairs_3d(s2, s2, rb, 1, period=60.0, weights1=w, weights2=w, num_threads=8)
        gm = hb.marked_npairs_3d(s2, s2, rb, 1, period=60.0, weights1=w, weights2=w)
        assert np.array_equal(gm, wm), (gm, wm)          # integer weights: exact
