"""Drop-in for ``halotools.mock_observables.pair_counters.npairs_projected``
(/root/reference/halotools/mock_observables/pair_counters/npairs_projected.py:20-227)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..helpers import array_is_monotonic, custom_len, check_num_threads_arg
from .mesh_helpers import _enclose_in_box, _set_approximate_cell_sizes, double_mesh_geometry

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
    c1 = _lib.Columns([x1in, y1in, z1in])
    c2 = c1 if (x2in is x1in and y2in is y1in and z2in is z1in) else _lib.Columns([x2in, y2in, z2in])
    g = geom.as_struct()
    rp = np.ascontiguousarray(rp_bins, dtype=np.float64)
    pi = np.array([pi_max], dtype=np.float64)
    _lib.run_engine(
        "htb_npairs_xy_z_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(rp), ctypes.c_int32(len(rp)), _lib._dp(pi), ctypes.c_int32(1),
        ctypes.c_int64(first), ctypes.c_int64(last),
        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), extra_flags=_lib.cache_flags(c1, c2, PBCs))
    return np.array(_dist.allreduce_sum(counts))


def _npairs_projected_process_args(sample1, sample2, rp_bins, pi_max, period,
                                   num_threads, approx_cell1_size, approx_cell2_size):
    """Same checks, defaults and error strings as npairs_projected.py:160-227."""
    num_threads = check_num_threads_arg(num_threads)

    same = sample2 is sample1
    x1 = sample1[:, 0]
    y1 = sample1[:, 1]
    z1 = sample1[:, 2]
    if same:
        x2, y2, z2 = x1, y1, z1
    else:
        x2 = sample2[:, 0]
        y2 = sample2[:, 1]
        z2 = sample2[:, 2]

    rp_bins = np.atleast_1d(rp_bins).astype('f8')
    try:
        assert rp_bins.ndim == 1
        assert len(rp_bins) > 1
        if len(rp_bins) > 2:
            assert array_is_monotonic(rp_bins, strict=True) == 1
    except AssertionError:
        msg = ("Input ``rp_bins`` must be a monotonically increasing 1D array "
               "with at least two entries")
        raise ValueError(msg)
    rp_max = np.max(rp_bins)

    if period is None:
        PBCs = False
        x1, y1, z1, x2, y2, z2, period = (
            _enclose_in_box(x1, y1, z1, x2, y2, z2,
                            min_size=[rp_max*3.0, rp_max*3.0, pi_max*3.0]))
    else:
        PBCs = True
        period = np.atleast_1d(period).astype(float)
        if len(period) == 1:
            period = np.array([period[0]]*3)
        try:
            assert np.all(period < np.inf)
            assert np.all(period > 0)
        except AssertionError:
            msg = "Input ``period`` must be a bounded positive number in all dimensions"
            raise ValueError(msg)

    try:
        assert pi_max > 0.
        assert pi_max < period[2]/3.
    except Exception:
        msg = ("Input ``pi_max`` must be a positive scalar less than period[2]/3")
        raise ValueError(msg)

    # NB the reference's default cell size is rp_max in all three dimensions here (npairs_projected.py:214-221)
    if approx_cell1_size is None:
        approx_cell1_size = [rp_max, rp_max, rp_max]
    elif custom_len(approx_cell1_size) == 1:
        approx_cell1_size = [approx_cell1_size, approx_cell1_size, approx_cell1_size]
    if approx_cell2_size is None:
        approx_cell2_size = [rp_max, rp_max, rp_max]
    elif custom_len(approx_cell2_size) == 1:
        approx_cell2_size = [approx_cell2_size, approx_cell2_size, approx_cell2_size]

    return (x1, y1, z1, x2, y2, z2,
            rp_bins, pi_max, period, num_threads, PBCs,
            approx_cell1_size, approx_cell2_size)
