"""Drop-in for ``halotools.mock_observables.pair_counters.npairs_s_mu``
(/root/reference/halotools/mock_observables/pair_counters/npairs_s_mu.py:20-209)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..helpers import array_is_monotonic
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry
from ._args import sample_columns
from .npairs_3d import _npairs_3d_process_args

__all__ = ("npairs_s_mu",)


def npairs_s_mu(sample1, sample2, s_bins, mu_bins, period=None, num_threads=1,
                approx_cell1_size=None, approx_cell2_size=None):
    """Pair counts in bins of redshift-space separation ``s`` and ``mu = cos(theta_LOS)`` (z is the
    line of sight); int64 (len(s_bins), len(mu_bins)), cumulative in s and in the engine's
    sin(theta_LOS) ordering, exactly as the reference returns them."""
    result = _npairs_3d_process_args(sample1, sample2, s_bins, period,
                                     num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    s_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]
    rmax = np.max(s_bins)

    mu_bins = np.atleast_1d(mu_bins)
    try:
        assert mu_bins.ndim == 1
        assert len(mu_bins) > 1
        if len(mu_bins) > 2:
            assert array_is_monotonic(mu_bins, strict=True) == 1
    except AssertionError:
        msg = ("\n Input `mu_bins` must be a monotonically increasing \n"
               "1D array with at least two entries")
        raise ValueError(msg)
    # the engine bins in sin(theta_LOS) (npairs_s_mu.py:174-175)
    mu_bins_prime = np.sort(np.sin(np.arccos(mu_bins)))

    search = [rmax, rmax, rmax]
    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)

    # every rank returns the 2-d cumulative sums of ITS differential histogram; cumulative sums are
    # linear, so the all-reduce of the per-rank results equals the single-GPU answer.
    counts = np.zeros((len(s_bins), len(mu_bins_prime)), dtype=np.int64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1, c2 = sample_columns([x1in, y1in, z1in], [x2in, y2in, z2in])
    g = geom.as_struct()
    sb = np.ascontiguousarray(s_bins, dtype=np.float64)
    mb = np.ascontiguousarray(mu_bins_prime, dtype=np.float64)
    _lib.run_engine(
        "htb_npairs_s_mu_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(sb), ctypes.c_int32(len(sb)), _lib._dp(mb), ctypes.c_int32(len(mb)),
        ctypes.c_int64(first), ctypes.c_int64(last),
        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), device=c1.device)
    return np.array(_dist.allreduce_sum(counts))
