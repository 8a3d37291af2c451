"""Drop-in for ``halotools.mock_observables.pair_counters.npairs_per_object_3d``
(/root/reference/halotools/mock_observables/pair_counters/npairs_per_object_3d.py:19-142)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry
from ._args import sample_columns
from .npairs_3d import _npairs_3d_process_args

__all__ = ("npairs_per_object_3d",)


def npairs_per_object_3d(sample1, sample2, rbins, period=None,
                         num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """counts[i, k] = number of ``sample2`` points within 3-d distance rbins[k] of sample1 point i;
    int64 (Npts1, len(rbins)), cumulative, rows in the input order of ``sample1``
    (cpairs/npairs_per_object_3d_engine.pyx:190-213)."""
    result = _npairs_3d_process_args(sample1, sample2, rbins, period,
                                     num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rbins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rmax = np.max(rbins)
    search = [rmax, rmax, rmax]
    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)

    c1, c2 = sample_columns([x1in, y1in, z1in], [x2in, y2in, z2in], host_only="npairs_per_object_3d")
    counts = np.zeros((c1.n, len(rbins)), dtype=np.int64)
    first, last = _dist.cell1_range(geom.ncells1)
    g = geom.as_struct()
    rb = np.ascontiguousarray(rbins, dtype=np.float64)
    _lib.run_engine(
        "htb_npairs_per_object_3d_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), extra_flags=_lib.cache_flags(c1, c2, PBCs))
    # rows of points outside this rank's mesh1 cells are zero: the sum over ranks is the full table
    # (npairs_per_object_3d.py:135-137)
    return _dist.allreduce_sum(counts)           # a fresh array already (no second copy of a large table)
