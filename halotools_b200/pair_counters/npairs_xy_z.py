"""Drop-in for ``halotools.mock_observables.pair_counters.npairs_xy_z``
(/root/reference/halotools/mock_observables/pair_counters/npairs_xy_z.py:20-163)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ._args import process_counter_args, sample_columns
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry

__all__ = ("npairs_xy_z",)


def npairs_xy_z(sample1, sample2, rp_bins, pi_bins, period=None, num_threads=1,
                approx_cell1_size=None, approx_cell2_size=None):
    """counts[k, g] = number of pairs with projected separation <= rp_bins[k] and line-of-sight
    (z) separation <= pi_bins[g]; int64 (len(rp_bins), len(pi_bins)), cumulative in both axes."""
    return _count(sample1, sample2, rp_bins, pi_bins, period, num_threads, approx_cell1_size, approx_cell2_size, None)


def _enqueue(out, sample1, sample2, rp_bins, pi_bins, period=None, num_threads=1,
             approx_cell1_size=None, approx_cell2_size=None):
    """The same count left ON THE DEVICE in ``out`` (int64 CUDA tensor, len(rp_bins) * len(pi_bins) entries): the call only
    enqueues work on the engine's stream (HTB_FLAG_DEVICE_OUTPUT); returns the objects to keep alive until the caller
    synchronises.  Multi-GPU: this rank's partial counts."""
    return _count(sample1, sample2, rp_bins, pi_bins, period, num_threads, approx_cell1_size, approx_cell2_size, out)


npairs_xy_z.enqueue = _enqueue


def _count(sample1, sample2, rp_bins, pi_bins, period, num_threads, approx_cell1_size, approx_cell2_size, out):
    result = _npairs_xy_z_process_args(sample1, sample2, rp_bins, pi_bins, period,
                                       num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rp_bins, pi_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rp_max = np.max(rp_bins)
    pi_max = np.max(pi_bins)
    search = [rp_max, rp_max, pi_max]
    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)

    counts = np.zeros((len(rp_bins), len(pi_bins)), dtype=np.int64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1, c2 = sample_columns([x1in, y1in, z1in], [x2in, y2in, z2in])
    g = geom.as_struct()
    rp = np.ascontiguousarray(rp_bins, dtype=np.float64)
    pi = np.ascontiguousarray(pi_bins, dtype=np.float64)
    _lib.run_engine(
        "htb_npairs_xy_z_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(rp), ctypes.c_int32(len(rp)), _lib._dp(pi), ctypes.c_int32(len(pi)),
        ctypes.c_int64(first), ctypes.c_int64(last),
        _lib.out_pointer(out, counts, ctypes.c_int64), device=c1.device,
        extra_flags=_lib.cache_flags(c1, c2, PBCs), out_device=out is not None)
    if out is not None:
        return (c1, c2)
    return np.array(_dist.allreduce_sum(counts))


def _npairs_xy_z_process_args(sample1, sample2, rp_bins, pi_bins, period,
                              num_threads, approx_cell1_size, approx_cell2_size):
    """The checks, defaults and error strings of npairs_xy_z.py:166-237 (shared processor: ``_args.py``)."""
    (c1, c2, (rp_bins, pi_bins), period, num_threads, PBCs,
     approx_cell1_size, approx_cell2_size) = process_counter_args(
        3, sample1, sample2, [(rp_bins, "rp_bins"), (pi_bins, "pi_bins")],
        lambda b: [np.max(b[0]), np.max(b[0]), np.max(b[1])],
        period, num_threads, approx_cell1_size, approx_cell2_size)
    return (c1[0], c1[1], c1[2], c2[0], c2[1], c2[2],
            rp_bins, pi_bins, period, num_threads, PBCs,
            approx_cell1_size, approx_cell2_size)
