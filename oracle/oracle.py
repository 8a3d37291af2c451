"""TEST INFRASTRUCTURE — Python face of the CPU oracle (oracle/pairs_oracle.c + oracle/mesh.py).

Engine-level restatement of the reference pair counters with the minimum of
front-end logic (defaults and the non-periodic enclosing box); argument
validation is NOT restated here — it lives in the product's host layer and is
tested against the reference's error strings directly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Parity status: PINNED (tests/test_oracle_*.py
check it against tests/golden/*.npz made by the unmodified reference, and, when
oracle/_ref exists, against the compiled reference engines).

Reference call sites restated:
  npairs_3d         /root/reference/halotools/mock_observables/pair_counters/npairs_3d.py:112-150,153-213
  npairs_xy_z       .../pair_counters/npairs_xy_z.py:122-163,166-237
  npairs_s_mu       .../pair_counters/npairs_s_mu.py:152-209
  marked_npairs_3d  .../pair_counters/marked_npairs_3d.py:130-183
  mean_delta_sigma  /root/reference/halotools/mock_observables/surface_density/mean_delta_sigma.py:212-255,258-330
  _enclose_in_box   .../pair_counters/mesh_helpers.py:17-64 ; _enclose_in_square :67-110
  npairs_projected      .../pair_counters/npairs_projected.py:118-157,160-227
  npairs_per_object_3d  .../pair_counters/npairs_per_object_3d.py:106-142
  marked_npairs_xy_z    .../pair_counters/marked_npairs_xy_z.py:141-194
  npairs_jackknife_3d / _xy_z  .../pair_counters/npairs_jackknife_3d.py:160-196, npairs_jackknife_xy_z.py:152-196
  weighted_npairs_xy    /root/reference/halotools/mock_observables/surface_density/weighted_npairs_xy.py:111-151,153-217
"""
import ctypes
import os
import subprocess

import numpy as np

from .mesh import DoubleMesh

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class _Geom(ctypes.Structure):
    _fields_ = [("ndivs1", ctypes.c_int * 3), ("ndivs2", ctypes.c_int * 3),
                ("cover", ctypes.c_int * 3), ("pbc", ctypes.c_int),
                ("period", ctypes.c_double * 3)]


def build(force=False):
    src = os.path.join(_HERE, "pairs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _geom(dm):
    g = _Geom()
    for d in range(3):
        if d < dm.ndim:
            g.ndivs1[d] = dm.mesh1.num_divs[d]
            g.ndivs2[d] = dm.mesh2.num_divs[d]
            g.cover[d] = dm.cover[d]
            g.period[d] = dm.period[d]
        else:
            g.ndivs1[d] = 1
            g.ndivs2[d] = 1
            g.cover[d] = 0
            g.period[d] = 1.0
    g.pbc = 1 if dm.PBCs else 0
    return g


def _triple(v, n=3):
    v = np.atleast_1d(np.asarray(v, dtype=float))
    if len(v) == 1:
        v = np.repeat(v, n)
    return v


def _enclose(cols1, cols2, min_size):
    lo = min(min(np.min(c) for c in cols1), min(np.min(c) for c in cols2))
    hi = max(max(np.max(c) for c in cols1), max(np.max(c) for c in cols2)) - lo
    cols1 = [c - lo for c in cols1]
    cols2 = [c - lo for c in cols2]
    L = np.array([hi] * len(cols1), dtype=float)
    if min_size is not None:
        ms = np.atleast_1d(np.asarray(min_size, dtype=float))
        L[L < ms] = ms[L < ms]
    return cols1, cols2, L


def build_double_mesh_3d(sample1, sample2, search, period, approx_cell1_size, approx_cell2_size):
    sample1 = np.asarray(sample1)
    sample2 = np.asarray(sample2)
    c1 = [sample1[:, d] for d in range(3)]
    c2 = [sample2[:, d] for d in range(3)]
    if period is None:
        pbc = False
        c1, c2, period = _enclose(c1, c2, [3.0 * s for s in search])
    else:
        pbc = True
        period = _triple(period)
    a1 = list(search) if approx_cell1_size is None else _triple(approx_cell1_size)
    a2 = list(search) if approx_cell2_size is None else _triple(approx_cell2_size)
    dm = DoubleMesh(c1, c2, a1, a2, search, period, pbc)
    return dm, c1, c2


def _sorted(cols, mesh):
    return [_f8(np.asarray(c)[mesh.idx_sorted]) for c in cols]


def _range(dm, cell1_range):
    if cell1_range is None:
        return 0, dm.mesh1.ncells
    first, last = int(cell1_range[0]), int(cell1_range[1])
    assert 0 <= first <= last <= dm.mesh1.ncells, "cell1_range %r outside the %d mesh1 cells" % (cell1_range, dm.mesh1.ncells)
    return first, last


def npairs_3d(sample1, sample2, rbins, period=None, approx_cell1_size=None, approx_cell2_size=None,
              num_threads=1, cell1_range=None, return_mesh=False):
    rbins = _f8(np.atleast_1d(rbins))
    rmax = float(np.max(rbins))
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, approx_cell1_size, approx_cell2_size)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    out = np.zeros(len(rbins), dtype=np.int64)
    lib().oracle_npairs_3d(ctypes.byref(g), _p(x1), _p(y1), _p(z1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                           _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                           _p(rbins), ctypes.c_int(len(rbins)), ctypes.c_int64(first), ctypes.c_int64(last),
                           ctypes.c_int(int(num_threads)), _p(out, ctypes.c_int64))
    return (out, dm) if return_mesh else out


def npairs_xy_z(sample1, sample2, rp_bins, pi_bins, period=None, approx_cell1_size=None,
                approx_cell2_size=None, num_threads=1, cell1_range=None, return_mesh=False):
    rp_bins = _f8(np.atleast_1d(rp_bins))
    pi_bins = _f8(np.atleast_1d(pi_bins))
    rp_max, pi_max = float(np.max(rp_bins)), float(np.max(pi_bins))
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, [rp_max, rp_max, pi_max], period,
                                      approx_cell1_size, approx_cell2_size)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    out = np.zeros((len(rp_bins), len(pi_bins)), dtype=np.int64)
    lib().oracle_npairs_xy_z(ctypes.byref(g), _p(x1), _p(y1), _p(z1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                             _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                             _p(rp_bins), ctypes.c_int(len(rp_bins)), _p(pi_bins), ctypes.c_int(len(pi_bins)),
                             ctypes.c_int64(first), ctypes.c_int64(last),
                             ctypes.c_int(int(num_threads)), _p(out, ctypes.c_int64))
    return (out, dm) if return_mesh else out


def npairs_s_mu(sample1, sample2, s_bins, mu_bins, period=None, approx_cell1_size=None,
                approx_cell2_size=None, num_threads=1, cell1_range=None):
    s_bins = _f8(np.atleast_1d(s_bins))
    rmax = float(np.max(s_bins))
    mu_prime = _f8(np.sort(np.sin(np.arccos(np.atleast_1d(mu_bins)))))  # npairs_s_mu.py:174-175
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, approx_cell1_size, approx_cell2_size)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    out = np.zeros((len(s_bins), len(mu_prime)), dtype=np.int64)
    lib().oracle_npairs_s_mu(ctypes.byref(g), _p(x1), _p(y1), _p(z1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                             _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                             _p(s_bins), ctypes.c_int(len(s_bins)), _p(mu_prime), ctypes.c_int(len(mu_prime)),
                             ctypes.c_int64(first), ctypes.c_int64(last),
                             ctypes.c_int(int(num_threads)), _p(out, ctypes.c_int64))
    return out


NUM_WEIGHTS = {1: 1, 2: 1, 3: 2, 4: 2, 5: 2, 6: 2, 7: 2, 8: 2, 9: 2, 10: 2, 11: 2,
               12: 4, 13: 4, 14: 3, 15: 3, 16: 5, 17: 5}  # marked_npairs_3d.py:281-326


def marked_npairs_3d(sample1, sample2, rbins, weight_func_id, period=None, weights1=None, weights2=None,
                     approx_cell1_size=None, approx_cell2_size=None, num_threads=1, cell1_range=None):
    rbins = _f8(np.atleast_1d(rbins))
    rmax = float(np.max(rbins))
    nw = NUM_WEIGHTS[int(weight_func_id)]
    n1, n2 = np.shape(sample1)[0], np.shape(sample2)[0]
    w1 = np.ones((n1, nw)) if weights1 is None else np.asarray(weights1, dtype=np.float64).reshape(n1, nw)
    w2 = np.ones((n2, nw)) if weights2 is None else np.asarray(weights2, dtype=np.float64).reshape(n2, nw)
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, approx_cell1_size, approx_cell2_size)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    w1s = _f8(w1[dm.mesh1.idx_sorted, :])
    w2s = _f8(w2[dm.mesh2.idx_sorted, :])
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    out = np.zeros(len(rbins), dtype=np.float64)
    lib().oracle_marked_npairs_3d(ctypes.byref(g), _p(x1), _p(y1), _p(z1),
                                  _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                                  _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                                  _p(w1s), _p(w2s), ctypes.c_int(nw), ctypes.c_int(int(weight_func_id)),
                                  _p(rbins), ctypes.c_int(len(rbins)), ctypes.c_int64(first), ctypes.c_int64(last),
                                  ctypes.c_int(int(num_threads)), _p(out))
    return out


def mean_delta_sigma(galaxies, particles, effective_particle_masses, rp_bins, period=None,
                     approx_cell1_size=None, approx_cell2_size=None, num_threads=1, per_object=False,
                     cell1_range=None, return_abs=False):
    """``return_abs``: also return A_ik, the sum of the absolute values of the terms accumulated into element
    (i, k) - the error scale of the SURVEY 8d per-object parity gate (per_object rows, input order)."""
    galaxies = np.asarray(galaxies, dtype=np.float64)
    particles = np.asarray(particles, dtype=np.float64)
    rp_bins = _f8(np.atleast_1d(rp_bins))
    rp_max = float(np.max(rp_bins))
    m = np.atleast_1d(np.asarray(effective_particle_masses, dtype=np.float64))
    if len(m) == 1:
        m = np.zeros(particles.shape[0]) + m[0]
    if period is None:
        # mean_delta_sigma.py:262-273 encloses both samples in a 3-D cube (no minimum size) and
        # keeps that cube's side as the period; the later ``if period is None`` branch (:303-307)
        # is therefore never taken.
        c1, c2, per3 = _enclose([galaxies[:, d] for d in range(3)], [particles[:, d] for d in range(3)], None)
        c1, c2, per2 = c1[:2], c2[:2], per3[:2]
        pbc = False
    else:
        c1 = [galaxies[:, 0], galaxies[:, 1]]
        c2 = [particles[:, 0], particles[:, 1]]
        per2 = _triple(period)[:2]
        pbc = True
    a1 = [rp_max] * 2 if approx_cell1_size is None else _triple(approx_cell1_size, 2)
    a2 = [rp_max] * 2 if approx_cell2_size is None else _triple(approx_cell2_size, 2)
    dm = DoubleMesh(c1, c2, a1, a2, [rp_max] * 2, per2, pbc)
    x1, y1 = _sorted(c1, dm.mesh1)
    x2, y2 = _sorted(c2, dm.mesh2)
    m2 = _f8(m[dm.mesh2.idx_sorted])
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    n1 = galaxies.shape[0]
    out = np.zeros((n1, len(rp_bins) - 1), dtype=np.float64)
    absout = np.zeros_like(out) if return_abs else None
    lib().oracle_mean_delta_sigma(ctypes.byref(g), _p(x1), _p(y1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                                  ctypes.c_int64(n1), _p(x2), _p(y2), _p(m2),
                                  _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                                  _p(rp_bins), ctypes.c_int(len(rp_bins)), ctypes.c_int64(first), ctypes.c_int64(last),
                                  ctypes.c_int(int(num_threads)), _p(out),
                                  _p(absout) if return_abs else ctypes.POINTER(ctypes.c_double)())
    unsort = np.empty(n1, dtype=np.int64)
    unsort[dm.mesh1.idx_sorted] = np.arange(n1)
    out = out[unsort, :]
    if return_abs:
        absout = absout[unsort, :]
        return (out, absout) if per_object else (np.mean(out, axis=0), np.mean(absout, axis=0))
    return out if per_object else np.mean(out, axis=0)


def npairs_projected(sample1, sample2, rp_bins, pi_max, period=None, approx_cell1_size=None,
                     approx_cell2_size=None, cell1_range=None):
    rp_bins = _f8(np.atleast_1d(rp_bins))
    rp_max, pi_max = float(np.max(rp_bins)), float(pi_max)
    # default cell sizes are rp_max in all three dimensions (npairs_projected.py:214-221)
    a1 = [rp_max] * 3 if approx_cell1_size is None else approx_cell1_size
    a2 = [rp_max] * 3 if approx_cell2_size is None else approx_cell2_size
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, [rp_max, rp_max, pi_max], period, a1, a2)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    out = np.zeros(len(rp_bins), dtype=np.int64)
    lib().oracle_npairs_projected(ctypes.byref(g), _p(x1), _p(y1), _p(z1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                                  _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                                  _p(rp_bins), ctypes.c_int(len(rp_bins)), ctypes.c_double(pi_max),
                                  ctypes.c_int64(first), ctypes.c_int64(last), _p(out, ctypes.c_int64))
    return out


def npairs_per_object_3d(sample1, sample2, rbins, period=None, approx_cell1_size=None, approx_cell2_size=None,
                         cell1_range=None):
    rbins = _f8(np.atleast_1d(rbins))
    rmax = float(np.max(rbins))
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, approx_cell1_size, approx_cell2_size)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    n1 = len(x1)
    out = np.zeros((n1, len(rbins)), dtype=np.int64)
    lib().oracle_npairs_per_object_3d(ctypes.byref(g), _p(x1), _p(y1), _p(z1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                                      ctypes.c_int64(n1),
                                      _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                                      _p(rbins), ctypes.c_int(len(rbins)),
                                      ctypes.c_int64(first), ctypes.c_int64(last), _p(out, ctypes.c_int64))
    unsort = np.empty(n1, dtype=np.int64)
    unsort[dm.mesh1.idx_sorted] = np.arange(n1)
    return out[unsort, :]


def marked_npairs_xy_z(sample1, sample2, rp_bins, pi_bins, period=None, weights1=None, weights2=None,
                       weight_func_id=0, approx_cell1_size=None, approx_cell2_size=None, cell1_range=None):
    rp_bins = _f8(np.atleast_1d(rp_bins))
    pi_bins = _f8(np.atleast_1d(pi_bins))
    rp_max, pi_max = float(np.max(rp_bins)), float(np.max(pi_bins))
    nw = NUM_WEIGHTS[int(weight_func_id)]
    n1, n2 = np.shape(sample1)[0], np.shape(sample2)[0]
    w1 = np.ones((n1, nw)) if weights1 is None else np.asarray(weights1, dtype=np.float64).reshape(n1, nw)
    w2 = np.ones((n2, nw)) if weights2 is None else np.asarray(weights2, dtype=np.float64).reshape(n2, nw)
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, [rp_max, rp_max, pi_max], period,
                                      approx_cell1_size, approx_cell2_size)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    w1s = _f8(w1[dm.mesh1.idx_sorted, :])
    w2s = _f8(w2[dm.mesh2.idx_sorted, :])
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    out = np.zeros((len(rp_bins), len(pi_bins)), dtype=np.float64)
    lib().oracle_marked_npairs_xy_z(ctypes.byref(g), _p(x1), _p(y1), _p(z1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                                    _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                                    _p(w1s), _p(w2s), ctypes.c_int(nw), ctypes.c_int(int(weight_func_id)),
                                    _p(rp_bins), ctypes.c_int(len(rp_bins)), _p(pi_bins), ctypes.c_int(len(pi_bins)),
                                    ctypes.c_int64(first), ctypes.c_int64(last), _p(out))
    return out


def weighted_npairs_xy(sample1, sample2, sample2_mass, rp_bins, period=None, approx_cell1_size=None,
                       approx_cell2_size=None, cell1_range=None, per_object=False):
    sample1 = np.asarray(sample1, dtype=np.float64)
    sample2 = np.asarray(sample2, dtype=np.float64)
    rp_bins = _f8(np.atleast_1d(rp_bins))
    rp_max = float(np.max(rp_bins))
    c1 = [sample1[:, 0], sample1[:, 1]]
    c2 = [sample2[:, 0], sample2[:, 1]]
    if period is None:
        c1, c2, per2 = _enclose(c1, c2, [3.0 * rp_max] * 2)       # weighted_npairs_xy.py:186-191
        pbc = False
    else:
        per2 = _triple(period, 2)[:2]
        pbc = True
    a1 = [rp_max] * 2 if approx_cell1_size is None else _triple(approx_cell1_size, 2)
    a2 = [rp_max] * 2 if approx_cell2_size is None else _triple(approx_cell2_size, 2)
    dm = DoubleMesh(c1, c2, a1, a2, [rp_max] * 2, per2, pbc)
    x1, y1 = _sorted(c1, dm.mesh1)
    x2, y2 = _sorted(c2, dm.mesh2)
    w2 = _f8(np.asarray(sample2_mass, dtype=np.float64)[dm.mesh2.idx_sorted])
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    if per_object:
        n1 = len(x1)
        rows = np.zeros((n1, len(rp_bins)), dtype=np.float64)
        lib().oracle_weighted_npairs_per_object_xy(
            ctypes.byref(g), _p(x1), _p(y1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64), ctypes.c_int64(n1),
            _p(x2), _p(y2), _p(w2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
            _p(rp_bins), ctypes.c_int(len(rp_bins)), ctypes.c_int64(first), ctypes.c_int64(last), _p(rows))
        unsort = np.empty(n1, dtype=np.int64)
        unsort[dm.mesh1.idx_sorted] = np.arange(n1)
        return rows[unsort, :]
    out = np.zeros(len(rp_bins), dtype=np.float64)
    lib().oracle_weighted_npairs_xy(ctypes.byref(g), _p(x1), _p(y1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                                    _p(x2), _p(y2), _p(w2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                                    _p(rp_bins), ctypes.c_int(len(rp_bins)), ctypes.c_int64(first), ctypes.c_int64(last),
                                    _p(out))
    return out


def weighted_npairs_per_object_xy(sample1, sample2, sample2_mass, rp_bins, period=None, approx_cell1_size=None,
                                  approx_cell2_size=None, cell1_range=None):
    """weighted_npairs_per_object_xy.py:113-152 (same front-end logic as weighted_npairs_xy)"""
    return weighted_npairs_xy(sample1, sample2, sample2_mass, rp_bins, period=period,
                              approx_cell1_size=approx_cell1_size, approx_cell2_size=approx_cell2_size,
                              cell1_range=cell1_range, per_object=True)


def total_mass_enclosed_per_cylinder(centers, particles, particle_masses, downsampling_factor, rp_bins, period,
                                     approx_cell1_size=None, approx_cell2_size=None):
    """mass_in_cylinders.py:211-231"""
    particles = np.asarray(particles, dtype=np.float64)
    m = np.atleast_1d(np.asarray(particle_masses, dtype=np.float64))
    if len(m) == 1:
        m = np.zeros(particles.shape[0]) + m[0]
    mean = np.mean(m)
    out = weighted_npairs_per_object_xy(centers, particles, m / mean, rp_bins, period=_triple(period)[:2],
                                        approx_cell1_size=approx_cell1_size, approx_cell2_size=approx_cell2_size)
    out *= downsampling_factor * mean
    return out


def _jackknife(sample1, sample2, bins0, bins1, jtags1, jtags2, N_samples, period, weights1, weights2,
               approx_cell1_size, approx_cell2_size, cell1_range):
    bins0 = _f8(np.atleast_1d(bins0))
    m0 = float(np.max(bins0))
    if bins1 is None:
        search, b1 = [m0] * 3, np.zeros(0)
    else:
        b1 = _f8(np.atleast_1d(bins1))
        search = [m0, m0, float(np.max(b1))]
    n1, n2 = np.shape(sample1)[0], np.shape(sample2)[0]
    w1 = np.ones(n1) if weights1 is None else np.asarray(weights1, dtype=np.float64)
    w2 = np.ones(n2) if weights2 is None else np.asarray(weights2, dtype=np.float64)
    dm, c1, c2 = build_double_mesh_3d(sample1, sample2, search, period, approx_cell1_size, approx_cell2_size)
    x1, y1, z1 = _sorted(c1, dm.mesh1)
    x2, y2, z2 = _sorted(c2, dm.mesh2)
    w1s, w2s = _f8(w1[dm.mesh1.idx_sorted]), _f8(w2[dm.mesh2.idx_sorted])
    t1 = np.ascontiguousarray(np.asarray(jtags1).astype(np.int64)[dm.mesh1.idx_sorted])
    t2 = np.ascontiguousarray(np.asarray(jtags2).astype(np.int64)[dm.mesh2.idx_sorted])
    g = _geom(dm)
    first, last = _range(dm, cell1_range)
    shape = (N_samples + 1, len(bins0)) if bins1 is None else (N_samples + 1, len(bins0), len(b1))
    out = np.zeros(shape, dtype=np.float64)
    lib().oracle_npairs_jackknife(ctypes.byref(g), _p(x1), _p(y1), _p(z1), _p(dm.mesh1.cell_id_indices, ctypes.c_int64),
                                  _p(x2), _p(y2), _p(z2), _p(dm.mesh2.cell_id_indices, ctypes.c_int64),
                                  _p(w1s), _p(w2s), _p(t1, ctypes.c_int64), _p(t2, ctypes.c_int64), ctypes.c_int(int(N_samples)),
                                  _p(bins0), ctypes.c_int(len(bins0)), _p(b1) if len(b1) else None, ctypes.c_int(len(b1)),
                                  ctypes.c_int64(first), ctypes.c_int64(last), _p(out))
    return out


def npairs_jackknife_3d(sample1, sample2, rbins, jtags1, jtags2, N_samples, period=None, weights1=None,
                        weights2=None, approx_cell1_size=None, approx_cell2_size=None, cell1_range=None):
    return _jackknife(sample1, sample2, rbins, None, jtags1, jtags2, N_samples, period, weights1, weights2,
                      approx_cell1_size, approx_cell2_size, cell1_range)


def npairs_jackknife_xy_z(sample1, sample2, rp_bins, pi_bins, jtags1, jtags2, N_samples, period=None, weights1=None,
                          weights2=None, approx_cell1_size=None, approx_cell2_size=None, cell1_range=None):
    return _jackknife(sample1, sample2, rp_bins, pi_bins, jtags1, jtags2, N_samples, period, weights1, weights2,
                      approx_cell1_size, approx_cell2_size, cell1_range)


def brute_npairs_3d(sample1, sample2, rbins, period=None):
    s1, s2, rbins = _f8(sample1), _f8(sample2), _f8(np.atleast_1d(rbins))
    out = np.zeros(len(rbins), dtype=np.int64)
    per = None if period is None else _f8(_triple(period))
    lib().oracle_brute_npairs_3d(_p(s1), ctypes.c_int64(len(s1)), _p(s2), ctypes.c_int64(len(s2)),
                                 _p(rbins), ctypes.c_int(len(rbins)),
                                 _p(per) if per is not None else None, _p(out, ctypes.c_int64))
    return out
