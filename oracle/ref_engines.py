"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference Cython engines in oracle/_ref.

oracle/_ref holds only shared objects compiled by oracle/build_ref.py from the
sources where they lie under /root/reference (nothing of the reference is in the
repo).  On the GPU box the reference python front-ends are not available, so the
engines are called the way the reference front-ends call them
(/root/reference/halotools/mock_observables/pair_counters/npairs_3d.py:135-148):
a double-mesh object, the unsorted coordinate arrays, the bins and a
``(first_cell1, last_cell1)`` tuple; work is fanned out over a
``multiprocessing.Pool`` and partial results are summed.  The double-mesh handed
to the engine is oracle/mesh.py's restatement wearing the attribute names the
engines read (npairs_3d_engine.pyx:47-96).

Used as (1) an extra check of the C oracle, (2) ``cpu_baseline.kind ==
"reference"`` / ``bench.py --impl reference``.
"""
import importlib
import multiprocessing
import os
import sys
from functools import partial

import numpy as np

from . import oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

_ENGINE_MODULES = {
    "npairs_3d_engine": "halotools.mock_observables.pair_counters.cpairs.npairs_3d_engine",
    "npairs_xy_z_engine": "halotools.mock_observables.pair_counters.cpairs.npairs_xy_z_engine",
    "npairs_s_mu_engine": "halotools.mock_observables.pair_counters.cpairs.npairs_s_mu_engine",
    "marked_npairs_3d_engine": "halotools.mock_observables.pair_counters.marked_cpairs.marked_npairs_3d_engine",
    "mean_delta_sigma_engine": "halotools.mock_observables.surface_density.engines.mean_delta_sigma_engine",
}


def available():
    return os.path.isdir(os.path.join(_REF, "halotools"))


def engine(name):
    if not available():
        raise RuntimeError("oracle/_ref is not built (run python oracle/build_ref.py where /root/reference exists)")
    if "halotools" in sys.modules and not any(
            os.path.abspath(p) == _REF for p in getattr(sys.modules["halotools"], "__path__", [])):
        # a full reference tree is already imported (golden generation); use its engines
        return getattr(importlib.import_module(_ENGINE_MODULES[name]), name)
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    return getattr(importlib.import_module(_ENGINE_MODULES[name]), name)


class _MeshView(object):
    def __init__(self, m):
        self.npts = m.npts
        self.ncells = m.ncells
        self.idx_sorted = m.idx_sorted
        self.cell_id_indices = m.cell_id_indices
        names = "xyz"
        for d in range(m.ndim):
            setattr(self, "num_%sdivs" % names[d], m.num_divs[d])
            setattr(self, "%scell_size" % names[d], m.cell_size[d])
            setattr(self, "%speriod" % names[d], m.period[d])


class _DoubleMeshView(object):
    def __init__(self, dm):
        self.mesh1 = _MeshView(dm.mesh1)
        self.mesh2 = _MeshView(dm.mesh2)
        self._PBCs = dm.PBCs
        names = "xyz"
        for d in range(dm.ndim):
            setattr(self, "%speriod" % names[d], dm.period[d])
            setattr(self, "search_%slength" % names[d], dm.search[d])
            setattr(self, "num_%scell2_per_%scell1" % (names[d], names[d]), dm.per[d])


def _cell1_tuples(ncells, num_threads):
    # pair_counters/mesh_helpers.py:183-221
    if num_threads == 1:
        return 1, [(0, ncells)]
    if num_threads > ncells:
        return ncells, [(a, a + 1) for a in range(ncells)]
    parts = [a for a in np.array_split(np.arange(ncells), num_threads) if len(a) > 0]
    return num_threads, [(int(a[0]), int(a[0]) + len(a)) for a in parts]


def _run(eng, ncells, num_threads, cell1_range=None):
    if cell1_range is not None:
        first, last = int(cell1_range[0]), int(cell1_range[1])
        # the engines do not check their cell1_tuple: a range past the mesh reads out of bounds
        assert 0 <= first <= last <= ncells, "cell1_range %r outside the %d mesh1 cells" % (cell1_range, ncells)
        if num_threads == 1:
            return eng((first, last))
        nt, tuples = _cell1_tuples(last - first, num_threads)
        tuples = [(a + first, b + first) for a, b in tuples]
    else:
        nt, tuples = _cell1_tuples(ncells, num_threads)
    if nt > 1:
        pool = multiprocessing.Pool(nt)
        try:
            result = pool.map(eng, tuples)
        finally:
            pool.close()
            pool.join()
        return np.sum(np.array(result), axis=0)
    return eng(tuples[0])


_WORKER_ENGINE = None       # set in the parent before the pool forks: the workers inherit the arrays (no pickling)


def _worker_call(cells):
    return _WORKER_ENGINE(cells)


class PreparedCount(object):
    """A reference npairs_3d call split into its phases so that a benchmark can time them apart: ``__init__`` builds
    the double mesh (what RectangularDoubleMesh does in the parent, npairs_3d.py:119-125) and forks the worker pool
    (:143); ``run(cell1_range)`` is the engine fan-out over the cells alone (:144-146).  The workers inherit the
    sample arrays through fork, so - unlike the reference's pool.map of a partial holding the arrays - nothing is
    pickled per task: the engine time is not diluted by set-up work when only a few cells are counted."""

    def __init__(self, sample1, sample2, rbins, period, num_threads):
        import time
        global _WORKER_ENGINE
        rbins = np.atleast_1d(rbins).astype("f8")
        rmax = float(np.max(rbins))
        t0 = time.perf_counter()
        dm, c1, c2 = _o.build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, None, None)
        self.seconds_mesh = time.perf_counter() - t0
        self.dm = dm
        self.ncells = dm.mesh1.ncells
        self.num_threads = int(num_threads)
        _WORKER_ENGINE = partial(engine("npairs_3d_engine"), _DoubleMeshView(dm), c1[0], c1[1], c1[2], c2[0], c2[1], c2[2], rbins)
        t0 = time.perf_counter()
        self.pool = multiprocessing.Pool(self.num_threads) if self.num_threads > 1 else None
        self.seconds_pool = time.perf_counter() - t0

    def run(self, cell1_range):
        first, last = int(cell1_range[0]), int(cell1_range[1])
        if self.pool is None:
            return np.array(_WORKER_ENGINE((first, last)))
        # one contiguous range per worker, the reference's own split (mesh_helpers.py:183-221); every engine call
        # re-gathers the sorted coordinates of both samples first (npairs_3d_engine.pyx:58-64) - part of the engine
        nt, tuples = _cell1_tuples(last - first, self.num_threads)
        tuples = [(a + first, b + first) for a, b in tuples]
        return np.sum(np.array(self.pool.map(_worker_call, tuples, chunksize=1)), axis=0)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None


def npairs_3d(sample1, sample2, rbins, period=None, approx_cell1_size=None, approx_cell2_size=None,
              num_threads=1, cell1_range=None):
    rbins = np.atleast_1d(rbins).astype("f8")
    rmax = float(np.max(rbins))
    dm, c1, c2 = _o.build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, approx_cell1_size, approx_cell2_size)
    eng = partial(engine("npairs_3d_engine"), _DoubleMeshView(dm), c1[0], c1[1], c1[2], c2[0], c2[1], c2[2], rbins)
    return np.array(_run(eng, dm.mesh1.ncells, num_threads, cell1_range))


def npairs_xy_z(sample1, sample2, rp_bins, pi_bins, period=None, approx_cell1_size=None,
                approx_cell2_size=None, num_threads=1, cell1_range=None):
    rp_bins = np.atleast_1d(rp_bins).astype("f8")
    pi_bins = np.atleast_1d(pi_bins).astype("f8")
    rp_max, pi_max = float(np.max(rp_bins)), float(np.max(pi_bins))
    dm, c1, c2 = _o.build_double_mesh_3d(sample1, sample2, [rp_max, rp_max, pi_max], period,
                                         approx_cell1_size, approx_cell2_size)
    eng = partial(engine("npairs_xy_z_engine"), _DoubleMeshView(dm), c1[0], c1[1], c1[2], c2[0], c2[1], c2[2],
                  rp_bins, pi_bins)
    return np.array(_run(eng, dm.mesh1.ncells, num_threads, cell1_range))


def npairs_s_mu(sample1, sample2, s_bins, mu_bins, period=None, approx_cell1_size=None,
                approx_cell2_size=None, num_threads=1):
    s_bins = np.atleast_1d(s_bins).astype("f8")
    rmax = float(np.max(s_bins))
    mu_prime = np.sort(np.sin(np.arccos(np.atleast_1d(mu_bins))))
    dm, c1, c2 = _o.build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, approx_cell1_size, approx_cell2_size)
    eng = partial(engine("npairs_s_mu_engine"), _DoubleMeshView(dm), c1[0], c1[1], c1[2], c2[0], c2[1], c2[2],
                  s_bins, mu_prime)
    return np.array(_run(eng, dm.mesh1.ncells, num_threads))


def marked_npairs_3d(sample1, sample2, rbins, weight_func_id, period=None, weights1=None, weights2=None,
                     approx_cell1_size=None, approx_cell2_size=None, num_threads=1, cell1_range=None):
    rbins = np.atleast_1d(rbins).astype("f8")
    rmax = float(np.max(rbins))
    nw = _o.NUM_WEIGHTS[int(weight_func_id)]
    n1, n2 = np.shape(sample1)[0], np.shape(sample2)[0]
    w1 = np.ones((n1, nw)) if weights1 is None else np.asarray(weights1, dtype=np.float64).reshape(n1, nw)
    w2 = np.ones((n2, nw)) if weights2 is None else np.asarray(weights2, dtype=np.float64).reshape(n2, nw)
    dm, c1, c2 = _o.build_double_mesh_3d(sample1, sample2, [rmax] * 3, period, approx_cell1_size, approx_cell2_size)
    eng = partial(engine("marked_npairs_3d_engine"), _DoubleMeshView(dm), c1[0], c1[1], c1[2], c2[0], c2[1], c2[2],
                  w1, w2, int(weight_func_id), rbins)
    return np.array(_run(eng, dm.mesh1.ncells, num_threads, cell1_range))


def mean_delta_sigma(galaxies, particles, effective_particle_masses, rp_bins, period, approx_cell1_size=None,
                     approx_cell2_size=None, num_threads=1, per_object=False, cell1_range=None):
    """The reference's own mean_delta_sigma_engine, driven as surface_density/mean_delta_sigma.py:214-255 drives
    it (periodic case): rows of galaxies outside ``cell1_range`` stay zero, as in each reference worker."""
    from .mesh import DoubleMesh
    galaxies = np.asarray(galaxies, dtype=np.float64)
    particles = np.asarray(particles, dtype=np.float64)
    rp_bins = np.atleast_1d(rp_bins).astype("f8")
    rp_max = float(np.max(rp_bins))
    m = np.atleast_1d(np.asarray(effective_particle_masses, dtype=np.float64))
    if len(m) == 1:
        m = np.zeros(particles.shape[0]) + m[0]
    per2 = [float(p) for p in (np.atleast_1d(period).tolist() * 2)[:2]]
    a1 = [rp_max] * 2 if approx_cell1_size is None else list(np.atleast_1d(approx_cell1_size).astype(float))[:2] * (2 if np.size(approx_cell1_size) == 1 else 1)
    a2 = [rp_max] * 2 if approx_cell2_size is None else list(np.atleast_1d(approx_cell2_size).astype(float))[:2] * (2 if np.size(approx_cell2_size) == 1 else 1)
    x1, y1 = np.ascontiguousarray(galaxies[:, 0]), np.ascontiguousarray(galaxies[:, 1])
    x2, y2 = np.ascontiguousarray(particles[:, 0]), np.ascontiguousarray(particles[:, 1])
    dm = DoubleMesh([x1, y1], [x2, y2], a1, a2, [rp_max] * 2, per2, True)
    eng = partial(engine("mean_delta_sigma_engine"), _DoubleMeshView(dm), x1, y1, x2, y2, m, rp_bins)
    out = np.array(_run(eng, dm.mesh1.ncells, num_threads, cell1_range))
    return out if per_object else np.mean(out, axis=0)
