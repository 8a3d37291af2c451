"""CPU tests (no GPU): pin the oracle (oracle/pairs_oracle.c + oracle/mesh.py) against the golden
vectors produced by the unmodified reference (tests/golden/make_golden.py), against the compiled
reference engines in oracle/_ref when they are present, and against O(N^2) brute force."""
import numpy as np
import pytest

from oracle import oracle, ref_engines
from tests.golden import cases

ENGINE_FUNCS = ("npairs_3d", "npairs_xy_z", "npairs_s_mu", "marked_npairs_3d", "mean_delta_sigma",
                "npairs_projected", "npairs_per_object_3d", "marked_npairs_xy_z", "weighted_npairs_xy",
                "weighted_npairs_per_object_xy", "total_mass_enclosed_per_cylinder",
                "npairs_jackknife_3d", "npairs_jackknife_xy_z")
ENGINE_CASES = [n for n in cases.names() if cases._cases()[n][0] in ENGINE_FUNCS and n != "n3d_c1_full"]


def run_oracle(name, **extra):
    fn, args, kwargs = cases.get(name)
    kwargs.update(extra)
    return cases.flatten(getattr(oracle, fn)(*args, **kwargs))


@pytest.mark.parametrize("name", ENGINE_CASES)
def test_oracle_matches_reference_golden(name, golden):
    fn = cases._cases()[name][0]
    got = run_oracle(name)
    want = golden(name)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape
        if fn in ("npairs_3d", "npairs_xy_z", "npairs_s_mu", "npairs_projected", "npairs_per_object_3d"):
            assert g.dtype == np.int64
            assert np.array_equal(g, w)
        elif fn in ("marked_npairs_3d", "marked_npairs_xy_z", "weighted_npairs_xy", "weighted_npairs_per_object_xy", "total_mass_enclosed_per_cylinder", "npairs_jackknife_3d",
                    "npairs_jackknife_xy_z"):
            assert np.allclose(g, w, rtol=1e-12, atol=0)
        else:
            # Delta Sigma is a cancelling difference: abs + rel tolerance
            scale = np.max(np.abs(w))
            assert np.allclose(g, w, rtol=1e-10, atol=1e-12 * scale)


def test_oracle_config1_known_answer(golden):
    want = np.array([100000, 100004, 100020, 100066, 100254, 100790, 102408, 107560, 123606, 173566,
                     328856, 812270, 2314792, 6988496, 21546088])
    assert np.array_equal(golden("n3d_c1_full")[0], want)
    got = run_oracle("n3d_c1_full", num_threads=8)[0]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["n3d_periodic", "n3d_nonperiodic", "n3d_noncubic", "n3d_cellsizes_b"])
def test_oracle_threads_and_cell_ranges_agree(name):
    serial = run_oracle(name)[0]
    assert np.array_equal(run_oracle(name, num_threads=3)[0], serial)
    fn, args, kwargs = cases.get(name)
    _, dm = oracle.npairs_3d(*args, return_mesh=True, **kwargs)
    nc = dm.mesh1.ncells
    parts = [run_oracle(name, cell1_range=(a, b))[0] for a, b in ((0, nc // 3), (nc // 3, nc // 2), (nc // 2, nc))]
    assert np.array_equal(sum(parts), serial)


@pytest.mark.parametrize("period", [1.0, None, [1.0, 2.0, 3.0]])
def test_oracle_vs_brute_force(period):
    s1, s2 = cases.pts(1, 300), cases.pts(2, 400)
    if isinstance(period, list):
        s1, s2 = s1 * np.array(period), s2 * np.array(period)
    rbins = np.array([0.0, 0.05, 0.11, 0.2, 0.3])
    assert np.array_equal(oracle.npairs_3d(s1, s2, rbins, period=period), oracle.brute_npairs_3d(s1, s2, rbins, period))


@pytest.mark.skipif(not ref_engines.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["n3d_periodic", "n3d_nonperiodic", "n3d_c1_small", "n3d_on_boundaries"])
def test_oracle_vs_compiled_reference_engine(name):
    fn, args, kwargs = cases.get(name)
    assert np.array_equal(ref_engines.npairs_3d(*args, **kwargs), oracle.npairs_3d(*args, **kwargs))
    assert np.array_equal(ref_engines.npairs_3d(*args, num_threads=2, **kwargs), oracle.npairs_3d(*args, **kwargs))


@pytest.mark.skipif(not ref_engines.available(), reason="oracle/_ref not built")
def test_compiled_reference_engines_other_variants():
    fn, args, kwargs = cases.get("xyz_periodic")
    assert np.array_equal(ref_engines.npairs_xy_z(*args, **kwargs), oracle.npairs_xy_z(*args, **kwargs))
    fn, args, kwargs = cases.get("smu_periodic")
    assert np.array_equal(ref_engines.npairs_s_mu(*args, **kwargs), oracle.npairs_s_mu(*args, **kwargs))
    fn, args, kwargs = cases.get("marked_id13")
    assert np.allclose(ref_engines.marked_npairs_3d(*args, **kwargs), oracle.marked_npairs_3d(*args, **kwargs),
                       rtol=1e-12)


def test_visited_pairs_config1():
    s = cases.pts(43, 100000, 250.0)
    _, dm = oracle.npairs_3d(s[:10], s[:10], np.logspace(-1, np.log10(20), 15), period=250.0, return_mesh=True)
    assert dm.mesh1.num_divs == [12, 12, 12] and dm.mesh2.num_divs == [12, 12, 12]
    dm2 = oracle.build_double_mesh_3d(s, s, [20.0] * 3, 250.0, None, None)[0]
    assert abs(dm2.visited_pairs() - 1.563e8) / 1.563e8 < 2e-3


# ---------------------------------------------------------------- O(N^2) numpy cross-checks of the 8(f) restatements
def _min_image(a, b, L):
    d = a[:, None, :] - b[None, :, :]
    if L is not None:
        d = d - L * np.round(d / L)
    return d


@pytest.mark.parametrize("period", [1.0, None])
def test_oracle_8f_counters_vs_brute_force(period):
    """independent of the mesh / window logic: all-pairs numpy with the minimum image (search lengths < L/3, so the
    mesh's image is the nearest one); integer results identical, float sums to 1e-12"""
    rng = np.random.RandomState(3)
    s1, s2 = rng.uniform(0, 1, (300, 3)), rng.uniform(0, 1, (400, 3))
    w1, w2 = rng.uniform(0.5, 1.5, 300), rng.uniform(0.5, 1.5, 400)
    t1, t2 = rng.randint(1, 5, 300), rng.randint(1, 5, 400)
    rb = np.array([0.0, 0.05, 0.11, 0.2, 0.3])
    pi = np.array([0.0, 0.07, 0.25])
    d = _min_image(s1, s2, period)
    dsq = d[..., 0] ** 2 + d[..., 1] ** 2 + d[..., 2] ** 2
    dxy, dz = d[..., 0] ** 2 + d[..., 1] ** 2, d[..., 2] ** 2
    in_r = dsq[..., None] <= (rb ** 2)[None, None, :]
    in_rp = dxy[..., None] <= (rb ** 2)[None, None, :]
    in_pi = dz[..., None] <= (pi ** 2)[None, None, :]
    # npairs_per_object_3d, npairs_projected
    assert np.array_equal(oracle.npairs_per_object_3d(s1, s2, rb, period=period), in_r.sum(axis=1))
    assert np.array_equal(oracle.npairs_projected(s1, s2, rb[1:], 0.25, period=period),
                          (in_rp[..., 1:] & in_pi[..., 2:3]).sum(axis=(0, 1)))
    # marked_npairs_xy_z with product marks
    ww = w1[:, None] * w2[None, :]
    want = np.einsum("ij,ijk,ijg->kg", ww, in_rp.astype(float), in_pi.astype(float))
    got = oracle.marked_npairs_xy_z(s1, s2, rb, pi, period=period, weights1=w1, weights2=w2, weight_func_id=1)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    # jackknife: jweight restated with masks (npairs_jackknife_3d_engine.pyx:283-289)
    got = oracle.npairs_jackknife_3d(s1, s2, rb, t1, t2, 4, period=period, weights1=w1, weights2=w2)
    for s in range(5):
        a, b = (t1 == s)[:, None], (t2 == s)[None, :]
        jw = ww if s == 0 else np.where(a & b, 0.0, np.where(a | b, 0.5 * ww, ww))
        assert np.allclose(got[s], np.einsum("ij,ijk->k", jw, in_r.astype(float)), rtol=1e-12, atol=1e-12)
    # 2-d weighted counters
    per2 = None if period is None else [1.0, 1.0]
    d2 = _min_image(s1[:, :2], s2[:, :2], period)
    in2 = (d2[..., 0] ** 2 + d2[..., 1] ** 2)[..., None] <= (rb ** 2)[None, None, :]
    rows = np.einsum("j,ijk->ik", w2, in2.astype(float))
    assert np.allclose(oracle.weighted_npairs_per_object_xy(s1[:, :2], s2[:, :2], w2, rb, period=per2), rows, rtol=1e-12, atol=0)
    assert np.allclose(oracle.weighted_npairs_xy(s1[:, :2], s2[:, :2], w2, rb, period=per2), rows.sum(axis=0), rtol=1e-12)
