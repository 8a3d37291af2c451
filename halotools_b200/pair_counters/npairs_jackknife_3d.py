"""Drop-ins for ``halotools.mock_observables.pair_counters.npairs_jackknife_3d`` and ``npairs_jackknife_xy_z``
(/root/reference/halotools/mock_observables/pair_counters/npairs_jackknife_3d.py:29-262,
npairs_jackknife_xy_z.py:23-262)."""
import ctypes
from warnings import warn

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..custom_exceptions import HalotoolsError
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry
from ._args import sample_columns
from .npairs_3d import _npairs_3d_process_args
from .npairs_xy_z import _npairs_xy_z_process_args

__all__ = ("npairs_jackknife_3d", "npairs_jackknife_xy_z")


def npairs_jackknife_3d(sample1, sample2, rbins, jtags1, jtags2, N_samples,
                        period=None, weights1=None, weights2=None, num_threads=1,
                        approx_cell1_size=None, approx_cell2_size=None):
    """Weighted pair counts for the full sample (row 0) and with each of the ``N_samples`` jackknife sub-volumes left
    out (rows 1..N_samples): float64 (N_samples+1, len(rbins)), cumulative in r.  A pair counts w1*w2 if neither point
    lies in the removed sub-volume, half of that if one does (cpairs/npairs_jackknife_3d_engine.pyx:237-291)."""
    result = _npairs_3d_process_args(sample1, sample2, rbins, period,
                                     num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rbins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rmax = np.max(rbins)
    search = [rmax, rmax, rmax]
    weights1, weights2, jtags1, jtags2 = _process_weights_jtags(sample1, sample2, weights1, weights2,
                                                                jtags1, jtags2, N_samples)
    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)
    rb = np.ascontiguousarray(rbins, dtype=np.float64)
    counts = np.zeros((N_samples + 1, len(rb)), dtype=np.float64)
    _run("htb_npairs_jackknife_3d_engine", geom, (x1in, y1in, z1in), (x2in, y2in, z2in), weights1, weights2,
         jtags1, jtags2, N_samples, [(_lib._dp(rb), ctypes.c_int32(len(rb)))], counts, keep=(rb,))
    return np.array(_dist.allreduce_sum(counts))


def npairs_jackknife_xy_z(sample1, sample2, rp_bins, pi_bins,
                          jtags1, jtags2, N_samples,
                          period=None, weights1=None, weights2=None, num_threads=1,
                          approx_cell1_size=None, approx_cell2_size=None):
    """The (rp, pi) version: float64 (N_samples+1, len(rp_bins), len(pi_bins))
    (cpairs/npairs_jackknife_xy_z_engine.pyx:222-246)."""
    result = _npairs_xy_z_process_args(sample1, sample2, rp_bins, pi_bins, period,
                                       num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rp_bins, pi_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rp_max = np.max(rp_bins)
    pi_max = np.max(pi_bins)
    search = [rp_max, rp_max, pi_max]
    weights1, weights2, jtags1, jtags2 = _process_weights_jtags(sample1, sample2, weights1, weights2,
                                                                jtags1, jtags2, N_samples)
    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)
    rp = np.ascontiguousarray(rp_bins, dtype=np.float64)
    pi = np.ascontiguousarray(pi_bins, dtype=np.float64)
    counts = np.zeros((N_samples + 1, len(rp), len(pi)), dtype=np.float64)
    _run("htb_npairs_jackknife_xy_z_engine", geom, (x1in, y1in, z1in), (x2in, y2in, z2in), weights1, weights2,
         jtags1, jtags2, N_samples,
         [(_lib._dp(rp), ctypes.c_int32(len(rp))), (_lib._dp(pi), ctypes.c_int32(len(pi)))], counts, keep=(rp, pi))
    return np.array(_dist.allreduce_sum(counts))


def _run(entry, geom, cols1, cols2, weights1, weights2, jtags1, jtags2, N_samples, bins, counts, keep):
    first, last = _dist.cell1_range(geom.ncells1)
    same = all(a is b for a, b in zip(cols1, cols2))
    c1, c2 = sample_columns(cols1, cols1 if same else cols2, host_only="the jackknife pair counters")
    w1 = np.ascontiguousarray(weights1, dtype=np.float64)
    w2 = np.ascontiguousarray(weights2, dtype=np.float64)
    t1 = np.ascontiguousarray(jtags1, dtype=np.int64)
    t2 = np.ascontiguousarray(jtags2, dtype=np.int64)
    g = geom.as_struct()
    args = [ctypes.byref(g),
            c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
            c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
            _lib._dp(w1), _lib._dp(w2),
            t1.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), t2.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
            ctypes.c_int32(int(N_samples))]
    for ptr, n in bins:
        args += [ptr, n]
    args += [ctypes.c_int64(first), ctypes.c_int64(last), _lib._dp(counts)]
    _lib.run_engine(entry, *args)


def _process_weights_jtags(sample1, sample2, weights1, weights2, jtags1, jtags2, N_samples):
    """Same checks, defaults, errors and warnings as npairs_jackknife_3d.py:199-262."""
    if weights1 is None:
        weights1 = np.ones(np.shape(sample1)[0], dtype=np.float64)
    else:
        weights1 = np.asarray(weights1).astype("float64")
        if np.shape(weights1)[0] != np.shape(sample1)[0]:
            raise HalotoolsError("weights1 should have same len as sample1")
    if weights2 is None:
        weights2 = np.ones(np.shape(sample2)[0], dtype=np.float64)
    else:
        weights2 = np.asarray(weights2).astype("float64")
        if np.shape(weights2)[0] != np.shape(sample2)[0]:
            raise HalotoolsError("weights2 should have same len as sample2")

    jtags1 = np.asarray(jtags1).astype("int")
    if np.shape(jtags1)[0] != np.shape(sample1)[0]:
        raise HalotoolsError("jtags1 should have same len as sample1")
    jtags2 = np.asarray(jtags2).astype("int")
    if np.shape(jtags2)[0] != np.shape(sample2)[0]:
        raise HalotoolsError("jtags2 should have same len as sample2")

    if np.min(jtags1) < 1:
        raise HalotoolsError("jtags1 must be >= 1")
    if np.min(jtags2) < 1:
        raise HalotoolsError("jtags2 must be >= 1")
    if np.max(jtags1) > N_samples:
        raise HalotoolsError("jtags1 must be <= N_samples")
    if np.max(jtags2) > N_samples:
        raise HalotoolsError("jtags2 must be <= N_samples")

    # the reference tests np.unique(jtags1) against 1..N_samples twice (npairs_jackknife_3d.py:256-259; the second
    # message names sample2); with the bounds checked above that is "every tag occurs", decided here by a bincount
    if not np.all(np.bincount(jtags1, minlength=N_samples + 1)[1:] > 0):
        warn("Warning: sample1 does not contain points in every jackknife sample.")
        warn("Warning: sample2 does not contain points in every jackknife sample.")

    return weights1, weights2, jtags1, jtags2
