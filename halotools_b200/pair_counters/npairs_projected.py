"""Drop-in for ``halotools.mock_observables.pair_counters.npairs_projected``
(/root/reference/halotools/mock_observables/pair_counters/npairs_projected.py:20-227)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ._args import process_counter_args, sample_columns
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry

__all__ = ("npairs_projected",)


def npairs_projected(sample1, sample2, rp_bins, pi_max, period=None,
                     num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """counts[k] = number of pairs with projected (xy) separation <= rp_bins[k] and line-of-sight (z)
    separation <= pi_max; int64 (len(rp_bins),), cumulative.  The reference's engine
    (cpairs/npairs_projected_engine.pyx:184-189) is the (rp, pi) counter with the single pi edge pi_max, so
    this front-end runs ``htb_npairs_xy_z_engine`` with one pi edge (the fast (rp, pi) kernel)."""
    result = _npairs_projected_process_args(sample1, sample2, rp_bins, pi_max, period,
                                            num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rp_bins, pi_max, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rp_max = np.max(rp_bins)
    search = [rp_max, rp_max, pi_max]
    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)

    counts = np.zeros(len(rp_bins), dtype=np.int64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1, c2 = sample_columns([x1in, y1in, z1in], [x2in, y2in, z2in])
    g = geom.as_struct()
    rp = np.ascontiguousarray(rp_bins, dtype=np.float64)
    pi = np.array([pi_max], dtype=np.float64)
    _lib.run_engine(
        "htb_npairs_xy_z_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(rp), ctypes.c_int32(len(rp)), _lib._dp(pi), ctypes.c_int32(1),
        ctypes.c_int64(first), ctypes.c_int64(last),
        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), device=c1.device,
        extra_flags=_lib.cache_flags(c1, c2, PBCs))
    return np.array(_dist.allreduce_sum(counts))


def _npairs_projected_process_args(sample1, sample2, rp_bins, pi_max, period,
                                   num_threads, approx_cell1_size, approx_cell2_size):
    """The checks, defaults and error strings of npairs_projected.py:160-227 (shared processor: ``_args.py``).
    NB the reference's default cell size is rp_max in all three dimensions here (:214-221)."""
    def after_period(checked, per):
        try:
            assert pi_max > 0.
            assert pi_max < per[2]/3.
        except Exception:
            msg = ("Input ``pi_max`` must be a positive scalar less than period[2]/3")
            raise ValueError(msg)

    (c1, c2, (rp_bins,), period, num_threads, PBCs,
     approx_cell1_size, approx_cell2_size) = process_counter_args(
        3, sample1, sample2, [(rp_bins, "rp_bins")],
        lambda b: [np.max(b[0]), np.max(b[0]), pi_max],
        period, num_threads, approx_cell1_size, approx_cell2_size,
        cell_default=lambda b: [np.max(b[0])] * 3, after_period=after_period)
    return (c1[0], c1[1], c1[2], c2[0], c2[1], c2[2],
            rp_bins, pi_max, period, num_threads, PBCs,
            approx_cell1_size, approx_cell2_size)
