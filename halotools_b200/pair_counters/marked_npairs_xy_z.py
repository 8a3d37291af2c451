"""Drop-in for ``halotools.mock_observables.pair_counters.marked_npairs_xy_z``
(/root/reference/halotools/mock_observables/pair_counters/marked_npairs_xy_z.py:23-335)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..custom_exceptions import HalotoolsError
from .marked_npairs_3d import _process_one_weights
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry
from ._args import sample_columns
from .npairs_xy_z import _npairs_xy_z_process_args

__all__ = ("marked_npairs_xy_z",)

# weights per point expected by each weight_func_id; this front-end of the reference knows ids 1-15 only and its
# default weight_func_id=0 is rejected (marked_npairs_xy_z.py:292-335)
_NUM_WEIGHTS = {1: 1, 2: 1, 3: 2, 4: 2, 5: 2, 6: 2, 7: 2, 8: 2, 9: 2, 10: 2, 11: 2,
                12: 4, 13: 4, 14: 3, 15: 3}


def marked_npairs_xy_z(sample1, sample2, rp_bins, pi_bins,
                       period=None, weights1=None, weights2=None,
                       weight_func_id=0, num_threads=1,
                       approx_cell1_size=None, approx_cell2_size=None):
    """Sum of f_id(w1_i, w2_j) over pairs with projected separation <= rp_bins[k] and line-of-sight separation
    <= pi_bins[g]; float64 (len(rp_bins), len(pi_bins)), cumulative in both axes
    (marked_cpairs/marked_npairs_xy_z_engine.pyx:209-225)."""
    result = _npairs_xy_z_process_args(sample1, sample2, rp_bins, pi_bins, period,
                                       num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rp_bins, pi_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rp_max = np.max(rp_bins)
    pi_max = np.max(pi_bins)
    search = [rp_max, rp_max, pi_max]

    weights1, weights2 = _marked_npairs_process_weights(sample1, sample2,
                                                        weights1, weights2, weight_func_id)

    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)

    counts = np.zeros((len(rp_bins), len(pi_bins)), dtype=np.float64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1, c2 = sample_columns([x1in, y1in, z1in], [x2in, y2in, z2in], host_only="marked_npairs_xy_z")
    w1 = np.ascontiguousarray(weights1, dtype=np.float64)
    w2 = w1 if weights2 is weights1 else np.ascontiguousarray(weights2, dtype=np.float64)
    g = geom.as_struct()
    rp = np.ascontiguousarray(rp_bins, dtype=np.float64)
    pi = np.ascontiguousarray(pi_bins, dtype=np.float64)
    _lib.run_engine(
        "htb_marked_npairs_xy_z_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(w1), _lib._dp(w2), ctypes.c_int32(w1.shape[1]), ctypes.c_int32(int(weight_func_id)),
        _lib._dp(rp), ctypes.c_int32(len(rp)), _lib._dp(pi), ctypes.c_int32(len(pi)),
        ctypes.c_int64(first), ctypes.c_int64(last),
        _lib._dp(counts), extra_flags=_lib.cache_flags(c1, c2, PBCs))
    return np.array(_dist.allreduce_sum(counts))


def _marked_npairs_process_weights(sample1, sample2, weights1, weights2, weight_func_id):
    """weights -> float64 (Npts, n_w) with n_w fixed by ``weight_func_id``; HalotoolsError on any
    shape mismatch (marked_npairs_xy_z.py:197-289)."""
    correct_num_weights = _func_signature_int_from_wfunc(weight_func_id)
    same = (weights2 is weights1) and (weights1 is not None) and (np.shape(sample1)[0] == np.shape(sample2)[0])
    weights1 = _process_one_weights(weights1, np.shape(sample1)[0], correct_num_weights, weight_func_id, 1)
    if same:
        return weights1, weights1
    weights2 = _process_one_weights(weights2, np.shape(sample2)[0], correct_num_weights, weight_func_id, 2)
    return weights1, weights2


def _func_signature_int_from_wfunc(weight_func_id):
    """Number of weights per point a weighting function reads (marked_npairs_xy_z.py:292-335)."""
    if type(weight_func_id) != int:
        msg = "\n weight_func_id parameter must be an integer ID of a weighting function."
        raise ValueError(msg)
    if weight_func_id in _NUM_WEIGHTS:
        return _NUM_WEIGHTS[weight_func_id]
    msg = ("The value ``weight_func_id`` = %i is not recognized")
    raise HalotoolsError(msg % weight_func_id)
