"""Drop-in for ``halotools.mock_observables.pair_counters.npairs_3d``
(/root/reference/halotools/mock_observables/pair_counters/npairs_3d.py:20-150): same
signature, same argument processing and errors; the mesh build and the pair loop run on the GPU
(csrc/mesh.cu + csrc/count.cu through htb_npairs_3d_engine)."""
import ctypes
import multiprocessing

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..helpers import array_is_monotonic, custom_len, check_num_threads_arg
from .mesh_helpers import _enclose_in_box, _set_approximate_cell_sizes, double_mesh_geometry

__all__ = ("npairs_3d",)


def npairs_3d(sample1, sample2, rbins, period=None, num_threads=1,
              approx_cell1_size=None, approx_cell2_size=None):
    """Number of pairs with 3-d separation <= each entry of ``rbins`` (cumulative), int64.

    Parameters follow the reference exactly (npairs_3d.py:22-74): ``sample1``/``sample2``
    (Npts, 3) arrays, ``rbins`` edges, ``period`` None / scalar / length-3, ``num_threads``
    (validated, then ignored: the GPU engine has no thread knob), ``approx_cell*_size``
    (honoured for the REFERENCE cell geometry that fixes which periodic shift a pair gets;
    it never changes the result).  If sample1 is sample2 pairs are double counted and
    every point pairs with itself, as in the reference.
    """
    result = _npairs_3d_process_args(sample1, sample2, rbins, period,
                                     num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rbins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rmax = np.max(rbins)
    search = [rmax, rmax, rmax]
    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)

    counts = np.zeros(len(rbins), dtype=np.int64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1 = _lib.Columns([x1in, y1in, z1in])
    c2 = c1 if (x2in is x1in and y2in is y1in and z2in is z1in) else _lib.Columns([x2in, y2in, z2in])
    g = geom.as_struct()
    rb = np.ascontiguousarray(rbins, dtype=np.float64)
    _lib.run_engine(
        "htb_npairs_3d_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), device=c1.device,
        extra_flags=_lib.cache_flags(c1, c2, PBCs))
    return np.array(_dist.allreduce_sum(counts))


def _npairs_3d_process_args(sample1, sample2, rbins, period,
                            num_threads, approx_cell1_size, approx_cell2_size):
    """Same checks, defaults and error strings as npairs_3d.py:153-213."""
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
    rbins = np.atleast_1d(rbins).astype('f8')
    rmax = np.max(rbins)

    try:
        assert rbins.ndim == 1
        assert len(rbins) > 1
        if len(rbins) > 2:
            assert array_is_monotonic(rbins, strict=True) == 1
    except AssertionError:
        msg = "Input ``rbins`` must be a monotonically increasing 1D array with at least two entries"
        raise ValueError(msg)

    if period is None:
        if getattr(sample1, "is_cuda", False) or getattr(sample2, "is_cuda", False):
            raise ValueError("device-resident samples need an explicit ``period``")
        PBCs = False
        x1, y1, z1, x2, y2, z2, period = (
            _enclose_in_box(x1, y1, z1, x2, y2, z2,
                            min_size=[rmax*3.0, rmax*3.0, rmax*3.0]))
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

    if approx_cell1_size is None:
        approx_cell1_size = [rmax, rmax, rmax]
    elif custom_len(approx_cell1_size) == 1:
        approx_cell1_size = [approx_cell1_size, approx_cell1_size, approx_cell1_size]
    if approx_cell2_size is None:
        approx_cell2_size = [rmax, rmax, rmax]
    elif custom_len(approx_cell2_size) == 1:
        approx_cell2_size = [approx_cell2_size, approx_cell2_size, approx_cell2_size]

    return (x1, y1, z1, x2, y2, z2,
            rbins, period, num_threads, PBCs,
            approx_cell1_size, approx_cell2_size)
