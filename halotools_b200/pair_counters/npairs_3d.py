"""Drop-in for ``halotools.mock_observables.pair_counters.npairs_3d``
(/root/reference/halotools/mock_observables/pair_counters/npairs_3d.py:20-150): same
signature, same argument processing and errors; the mesh build and the pair loop run on the GPU
(csrc/mesh.cu + csrc/count.cu through htb_npairs_3d_engine)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ._args import process_counter_args, sample_columns
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry

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
    return _count(sample1, sample2, rbins, period, num_threads, approx_cell1_size, approx_cell2_size, None)


def _enqueue(out, sample1, sample2, rbins, period=None, num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """The same count left ON THE DEVICE: ``out`` is an int64 CUDA tensor of len(rbins) entries; the call only enqueues
    work on the engine's stream (HTB_FLAG_DEVICE_OUTPUT) and returns the objects that must stay alive until the
    caller synchronises.  Multi-GPU: ``out`` holds this rank's partial counts (the caller all-reduces the table)."""
    return _count(sample1, sample2, rbins, period, num_threads, approx_cell1_size, approx_cell2_size, out)


npairs_3d.enqueue = _enqueue


def _count(sample1, sample2, rbins, period, num_threads, approx_cell1_size, approx_cell2_size, out):
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
    c1, c2 = sample_columns([x1in, y1in, z1in], [x2in, y2in, z2in])
    g = geom.as_struct()
    rb = np.ascontiguousarray(rbins, dtype=np.float64)
    _lib.run_engine(
        "htb_npairs_3d_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
        _lib.out_pointer(out, counts, ctypes.c_int64), device=c1.device,
        extra_flags=_lib.cache_flags(c1, c2, PBCs), out_device=out is not None)
    if out is not None:
        return (c1, c2)
    return np.array(_dist.allreduce_sum(counts))


def _npairs_3d_process_args(sample1, sample2, rbins, period,
                            num_threads, approx_cell1_size, approx_cell2_size):
    """The checks, defaults and error strings of npairs_3d.py:153-213 (shared processor: ``_args.py``)."""
    (c1, c2, (rbins,), period, num_threads, PBCs,
     approx_cell1_size, approx_cell2_size) = process_counter_args(
        3, sample1, sample2, [(rbins, "rbins")], lambda b: [np.max(b[0])] * 3,
        period, num_threads, approx_cell1_size, approx_cell2_size)
    return (c1[0], c1[1], c1[2], c2[0], c2[1], c2[2],
            rbins, period, num_threads, PBCs,
            approx_cell1_size, approx_cell2_size)
