"""Drop-in for ``halotools.mock_observables.surface_density.weighted_npairs_xy``
(/root/reference/halotools/mock_observables/surface_density/weighted_npairs_xy.py:22-217)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..pair_counters._args import process_counter_args, sample_columns
from ..pair_counters.mesh_helpers import _set_approximate_2d_cell_sizes, double_mesh_geometry

__all__ = ("weighted_npairs_xy",)


def weighted_npairs_xy(sample1, sample2, sample2_mass, rp_bins,
                       period=None, num_threads=1,
                       approx_cell1_size=None, approx_cell2_size=None):
    """Total mass of ``sample2`` inside z-aligned cylinders of radii ``rp_bins`` centred on the points of
    ``sample1``, summed over ``sample1``: float64 (len(rp_bins),), cumulative
    (surface_density/engines/weighted_npairs_xy_engine.pyx:150-175)."""
    result = _weighted_npairs_xy_process_args(sample1, sample2, sample2_mass,
                                              rp_bins, period, num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, x2in, y2in, w2in = result[0:5]
    rp_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[5:]

    rp_max = np.max(rp_bins)
    search = [rp_max, rp_max]
    approx_cell1_size, approx_cell2_size = _set_approximate_2d_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(2, approx_cell1_size, approx_cell2_size, search, period[:2], PBCs)

    counts = np.zeros(len(rp_bins), dtype=np.float64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1, c2 = sample_columns([x1in, y1in], [x2in, y2in], host_only="weighted_npairs_xy")
    w2 = np.ascontiguousarray(w2in, dtype=np.float64)
    g = geom.as_struct()
    rb = np.ascontiguousarray(rp_bins, dtype=np.float64)
    _lib.run_engine(
        "htb_weighted_npairs_xy_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(w2), _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
        _lib._dp(counts))
    return np.array(_dist.allreduce_sum(counts))


def _weighted_npairs_xy_process_args(sample1, sample2, w2, rp_bins, period,
                                     num_threads, approx_cell1_size, approx_cell2_size):
    """The checks, defaults and error strings of weighted_npairs_xy.py:153-217 (shared processor:
    ``pair_counters/_args.py``, two mesh dimensions)."""
    assert w2.shape[0] == sample2.shape[0]
    (c1, c2, (rp_bins,), period, num_threads, PBCs,
     approx_cell1_size, approx_cell2_size) = process_counter_args(
        2, sample1, sample2, [(rp_bins, "rp_bins")], lambda b: [np.max(b[0])] * 2,
        period, num_threads, approx_cell1_size, approx_cell2_size)
    return (c1[0], c1[1], c2[0], c2[1], w2, rp_bins, period, num_threads, PBCs,
            approx_cell1_size, approx_cell2_size)
