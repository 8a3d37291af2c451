"""Reference parity AT THE SIZES OF BASELINE.json (pytest -m gpu): the CUDA path against the UNMODIFIED reference
engines (oracle/_ref: the reference's own .pyx compiled from /root/reference by oracle/build_ref.py; the C oracle
port is the fall-back checker when _ref is absent) on configs 2, 3, 4 and 5.

A full reference run of configs 2/4/5 takes tens of minutes on the host cores, so - except where it is cheap
(config 3 in full, config 2's DD in full) - both sides count the SAME contiguous ranges of reference mesh1 cells:
the ``cell1_tuple`` every reference engine takes (npairs_3d_engine.pyx:17,103; marked_npairs_3d_engine.pyx:22;
mean_delta_sigma_engine.pyx:18) and the C ABI's ``first_cell1, last_cell1``.  The GPU call still sorts and meshes
the FULL samples with the refinement, tile list and redo queue of the full-size call; only the sample1 tiles are
restricted.  Ranges include the first and the last cells so that periodic wraps in every dimension are covered.

Integer counts: np.array_equal.  marked sums: rtol 1e-12.  delta-sigma per object: the SURVEY 8d gate
|gpu - ref| <= 1e-12 |ref| + 1e-12 A_ik (A_ik = sum of |terms| of element (i, k), from the C oracle).
"""
import os

import numpy as np
import pytest

import halotools_b200 as hb
from halotools_b200 import _lib, distributed, synthetic
from oracle import oracle, ref_engines

pytestmark = pytest.mark.gpu

CORES = max(1, min(os.cpu_count() or 1, 32))
REF = ref_engines if ref_engines.available() else oracle      # the reference's own engines when they travelled here


def ref_kind():
    return "reference" if REF is ref_engines else "port"


# ------------------------------------------------------------------ config 2: tpcf counts (the bench step)
@pytest.fixture(scope="module")
def c2():
    gal = synthetic.fakesim_zheng07_mock(560, 250.0, seed=43)
    ran = synthetic.uniform_points(44, 5_000_000, 250.0)
    return gal, ran, synthetic.config_rbins()


def test_config2_DD_full(c2):
    gal, _, rbins = c2
    want = REF.npairs_3d(gal, gal, rbins, period=250.0, num_threads=CORES)
    got = hb.npairs_3d(gal, gal, rbins, period=250.0)
    assert _lib.last_stats["path"] == 1
    assert np.array_equal(got, want), (got, want)


# mesh1 of config 2 has 12^3 = 1728 cells; first cells, an interior block, and the last cells (wraps in x, y, z)
C2_RANGES = [(0, 40), (850, 880), (1700, 1728)]


@pytest.mark.parametrize("cells", C2_RANGES)
def test_config2_DR_cells(c2, cells):
    gal, ran, rbins = c2
    want = REF.npairs_3d(gal, ran, rbins, period=250.0, num_threads=CORES, cell1_range=cells)
    with distributed.cell_range(*cells):
        got = hb.npairs_3d(gal, ran, rbins, period=250.0)
    assert np.array_equal(got, want), (got, want)


@pytest.mark.parametrize("cells", [(0, 16), (1712, 1728)])
def test_config2_RR_cells_device_resident(c2, cells):
    """The bench's dominant launch (symmetric Fast3 on the 5e6 randoms, device-resident input as in bench.py)."""
    import torch
    _, ran, rbins = c2
    want = REF.npairs_3d(ran, ran, rbins, period=250.0, num_threads=CORES, cell1_range=cells)
    ran_d = torch.from_numpy(ran).cuda()
    with distributed.cell_range(*cells):
        got = hb.npairs_3d(ran_d, ran_d, rbins, period=250.0)
    st = dict(_lib.last_stats)
    assert st["path"] == 1
    assert np.array_equal(got, want), (got, want)
    # the same cells without the symmetric shortcut and without culling: the reference's own visit list
    old = _lib.default_flags
    _lib.default_flags = _lib.FLAG_NO_SYM | _lib.FLAG_NO_CULL
    try:
        with distributed.cell_range(*cells):
            got2 = hb.npairs_3d(ran_d, ran_d, rbins, period=250.0)
    finally:
        _lib.default_flags = old
    assert np.array_equal(got2, want)


def test_config2_RR_full_is_sum_of_cell_ranges(c2):
    """Size-independent property at full size: the full RR count equals the sum over a partition of the cells
    (two of whose parts are pinned to the reference above)."""
    import torch
    _, ran, rbins = c2
    ran_d = torch.from_numpy(ran).cuda()
    full = hb.npairs_3d(ran_d, ran_d, rbins, period=250.0)
    total = np.zeros_like(full)
    for cells in ((0, 16), (16, 900), (900, 1712), (1712, 1728)):
        with distributed.cell_range(*cells):
            total += hb.npairs_3d(ran_d, ran_d, rbins, period=250.0)
    assert np.array_equal(full, total)
    assert full[0] >= len(ran)                      # every point pairs with itself


# ------------------------------------------------------------------ config 3: wp counts, in full
def test_config3_npairs_xy_z_full():
    s = synthetic.uniform_points(43, 2_000_000, 1000.0)
    rp = np.logspace(-1, np.log10(30), 15)
    pi = [0.0, 60.0]
    want = REF.npairs_xy_z(s, s, rp, pi, period=1000.0, num_threads=CORES)
    got = hb.npairs_xy_z(s, s, rp, pi, period=1000.0)
    assert _lib.last_stats["path"] == 1             # FastXYZ
    assert np.array_equal(got, want), (got, want)
    # rp_pi_tpcf's counter (BinQ: many pi edges) at the same size on a range of cells
    pim = np.linspace(0.0, 60.0, 13)
    cells = (0, 600)
    want = REF.npairs_xy_z(s, s, rp, pim, period=1000.0, num_threads=CORES, cell1_range=cells)
    with distributed.cell_range(*cells):
        got = hb.npairs_xy_z(s, s, rp, pim, period=1000.0)
    assert _lib.last_stats["path"] == 3
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ config 4: marked counts, 1e7 points
@pytest.fixture(scope="module")
def c4():
    rng = np.random.RandomState(43)
    s = rng.uniform(0, 1000.0, (10_000_000, 3))
    w = rng.uniform(0.5, 1.5, 10_000_000)
    return s, w, synthetic.config_rbins()


def _ncells1_3d(rmax, period):
    """Number of reference mesh1 cells of a default-cell-size call (oracle mesh restatement on two points)."""
    pts = np.array([[0.1, 0.1, 0.1], [0.2, 0.2, 0.2]])
    return oracle.build_double_mesh_3d(pts, pts, [rmax] * 3, period, None, None)[0].mesh1.ncells


# mesh1 of config 4: rmax = 10**log10(20) is a hair above 20, so 49^3 = 117649 cells (not 50^3).  The first 1200 cells
# (x-layer 0: wraps in x), an interior block, and the last 600 cells (wraps in x, y and z)
@pytest.mark.parametrize("where", ["first", "interior", "last"])
def test_config4_marked_and_unmarked_cells(c4, where):
    s, w, rbins = c4
    nc = _ncells1_3d(float(rbins.max()), 1000.0)
    assert nc == 49 ** 3
    cells = {"first": (0, 1200), "interior": (nc // 2, nc // 2 + 600), "last": (nc - 600, nc)}[where]
    want = REF.marked_npairs_3d(s, s, rbins, 1, period=1000.0, weights1=w, weights2=w, num_threads=CORES, cell1_range=cells)
    wantn = REF.npairs_3d(s, s, rbins, period=1000.0, num_threads=CORES, cell1_range=cells)
    assert np.all(np.isfinite(want)) and wantn[0] > 0
    with distributed.cell_range(*cells):
        got = hb.marked_npairs_3d(s, s, rbins, 1, period=1000.0, weights1=w, weights2=w)
        assert _lib.last_stats["path"] == 1          # MarkedQ
        gotn = hb.npairs_3d(s, s, rbins, period=1000.0)
    assert np.array_equal(gotn, wantn), (gotn, wantn)
    err = np.max(np.abs(got - want) / np.abs(want))
    assert err <= 1e-12, (err, got, want)


def test_config4_symmetric_partition_sums_to_reference_cells(c4):
    """HTB_FLAG_PARTITION_SUM (what the multi-GPU front-ends pass): per-part counts of the symmetric kernels differ
    from the reference's per-range counts, their sum over a partition does not.  Checked on the unmarked counter:
    two parts of a partition of ALL cells = the full count, whose first cells are pinned to the reference above."""
    s, w, rbins = c4
    nc = _ncells1_3d(float(rbins.max()), 1000.0)
    full = hb.npairs_3d(s, s, rbins, period=1000.0)
    old = _lib.default_flags
    total = np.zeros_like(full)
    try:
        _lib.default_flags = old | _lib.FLAG_PARTITION_SUM
        for cells in ((0, nc // 3), (nc // 3, nc)):
            with distributed.cell_range(*cells):
                total += hb.npairs_3d(s, s, rbins, period=1000.0)
    finally:
        _lib.default_flags = old
    assert np.array_equal(total, full)


# ------------------------------------------------------------------ config 5: delta-sigma, 1e6 x 1e8
def test_config5_delta_sigma_cells():
    ngal = int(os.environ.get("HTB_C5_NGAL", 1_000_000))
    nptcl = int(os.environ.get("HTB_C5_NPTCL", 100_000_000))
    gal = synthetic.uniform_points(43, ngal, 1000.0)
    ptcl = synthetic.uniform_points(44, nptcl, 1000.0)
    rp = np.logspace(-1, np.log10(30), 15)
    # mesh1: 33 x 33 cells.  The reference's own engine on 4 cells (first two: wraps in x and y; last two), one process each
    ranges = [(0, 2), (1087, 1089)]
    rows = {}
    if ref_engines.available():
        for cells in ranges:
            rows[cells] = ref_engines.mean_delta_sigma(gal, ptcl, 1.0, rp, 1000.0, num_threads=2, per_object=True,
                                                       cell1_range=cells)
    # the C oracle on 16 cells around them, with the A_ik accumulators of the per-object gate
    wide = [(0, 8), (1081, 1089)]
    for cells in wide:
        with distributed.cell_range(*cells):
            got = hb.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0, per_object=True)
        assert _lib.last_stats["path"] == 2          # DSigmaR, the config-5 kernel
        want, A = oracle.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0, num_threads=CORES, per_object=True,
                                          cell1_range=cells, return_abs=True)
        inside = A[:, -1] > 0
        assert 0 < inside.sum() < ngal
        assert not got[~inside].any()
        excess = np.abs(got - want) - (1e-12 * np.abs(want) + 1e-12 * A)
        assert np.max(excess) <= 0.0, (np.max(np.abs(got - want) / np.maximum(A, 1e-300)))
        for sub in ranges:
            if sub in rows and cells[0] <= sub[0] and sub[1] <= cells[1]:
                r = rows[sub]
                sel = r.any(axis=1)
                assert sel.sum() > 0
                ex = np.abs(got[sel] - r[sel]) - (1e-12 * np.abs(r[sel]) + 1e-12 * A[sel])
                assert np.max(ex) <= 0.0
        # column sums over these galaxies (what per_object=False returns, times N): 1e-12 of the summed |terms|
        with distributed.cell_range(*cells):
            mean = hb.mean_delta_sigma(gal, ptcl, 1.0, rp, period=1000.0)
        assert np.all(np.abs(mean * ngal - want.sum(axis=0)) <= 1e-12 * A.sum(axis=0))
