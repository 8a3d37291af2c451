"""Drop-in for ``halotools.mock_observables.surface_density.weighted_npairs_xy``
(/root/reference/halotools/mock_observables/surface_density/weighted_npairs_xy.py:22-217)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..helpers import array_is_monotonic, custom_len, check_num_threads_arg
from ..pair_counters.mesh_helpers import (_enclose_in_square, _set_approximate_2d_cell_sizes,
                                          double_mesh_geometry)

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
    c1 = _lib.Columns([x1in, y1in])
    c2 = _lib.Columns([x2in, y2in])
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
    """Same checks, defaults and error strings as weighted_npairs_xy.py:153-217."""
    num_threads = check_num_threads_arg(num_threads)

    x1 = sample1[:, 0]
    y1 = sample1[:, 1]
    x2 = sample2[:, 0]
    y2 = sample2[:, 1]

    assert w2.shape[0] == sample2.shape[0]

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
        x1, y1, x2, y2, period = (
            _enclose_in_square(x1, y1, x2, y2,
                               min_size=[rp_max*3.0, rp_max*3.0]))
    else:
        PBCs = True
        period = np.atleast_1d(period).astype(float)
        if len(period) == 1:
            period = np.array([period[0]]*2)
        try:
            assert np.all(period < np.inf)
            assert np.all(period > 0)
        except AssertionError:
            msg = "Input ``period`` must be a bounded positive number in all dimensions"
            raise ValueError(msg)

    if approx_cell1_size is None:
        approx_cell1_size = [rp_max, rp_max]
    elif custom_len(approx_cell1_size) == 1:
        approx_cell1_size = [approx_cell1_size, approx_cell1_size]
    if approx_cell2_size is None:
        approx_cell2_size = [rp_max, rp_max]
    elif custom_len(approx_cell2_size) == 1:
        approx_cell2_size = [approx_cell2_size, approx_cell2_size]

    return (x1, y1, x2, y2, w2, rp_bins, period, num_threads, PBCs,
            approx_cell1_size, approx_cell2_size)
