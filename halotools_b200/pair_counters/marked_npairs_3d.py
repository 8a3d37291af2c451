"""Drop-in for ``halotools.mock_observables.pair_counters.marked_npairs_3d``
(/root/reference/halotools/mock_observables/pair_counters/marked_npairs_3d.py:25-183)."""
import ctypes

import numpy as np

from .. import _lib
from .. import distributed as _dist
from ..custom_exceptions import HalotoolsError
from .mesh_helpers import _set_approximate_cell_sizes, double_mesh_geometry
from ._args import sample_columns
from .npairs_3d import _npairs_3d_process_args

__all__ = ("marked_npairs_3d",)

# weights per point expected by each weight_func_id (marked_npairs_3d.py:281-326)
_NUM_WEIGHTS = {1: 1, 2: 1, 3: 2, 4: 2, 5: 2, 6: 2, 7: 2, 8: 2, 9: 2, 10: 2, 11: 2,
                12: 4, 13: 4, 14: 3, 15: 3, 16: 5, 17: 5}


def marked_npairs_3d(sample1, sample2, rbins, weight_func_id, period=None,
                     weights1=None, weights2=None,
                     num_threads=1, approx_cell1_size=None, approx_cell2_size=None):
    """Sum of f_id(w1_i, w2_j) over pairs with 3-d separation <= each entry of ``rbins``
    (cumulative), float64.  The 17 weighting functions are the reference's
    (marked_cpairs/marking_functions.pyx:14-217)."""
    result = _npairs_3d_process_args(sample1, sample2, rbins, period,
                                     num_threads, approx_cell1_size, approx_cell2_size)
    x1in, y1in, z1in, x2in, y2in, z2in = result[0:6]
    rbins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size = result[6:]

    rmax = np.max(rbins)
    search = [rmax, rmax, rmax]

    weights1, weights2 = _marked_npairs_process_weights(sample1, sample2,
                                                        weights1, weights2, weight_func_id)

    approx_cell1_size, approx_cell2_size = _set_approximate_cell_sizes(
        approx_cell1_size, approx_cell2_size, period)
    geom = double_mesh_geometry(3, approx_cell1_size, approx_cell2_size, search, period, PBCs)

    counts = np.zeros(len(rbins), dtype=np.float64)
    first, last = _dist.cell1_range(geom.ncells1)
    c1, c2 = sample_columns([x1in, y1in, z1in], [x2in, y2in, z2in], host_only="marked_npairs_3d")
    w1 = np.ascontiguousarray(weights1, dtype=np.float64)
    w2 = w1 if weights2 is weights1 else np.ascontiguousarray(weights2, dtype=np.float64)
    g = geom.as_struct()
    rb = np.ascontiguousarray(rbins, dtype=np.float64)
    _lib.run_engine(
        "htb_marked_npairs_3d_engine", ctypes.byref(g),
        c1.ptrs[0], c1.ptrs[1], c1.ptrs[2], ctypes.c_int64(c1.stride), ctypes.c_int64(c1.n),
        c2.ptrs[0], c2.ptrs[1], c2.ptrs[2], ctypes.c_int64(c2.stride), ctypes.c_int64(c2.n),
        _lib._dp(w1), _lib._dp(w2), ctypes.c_int32(w1.shape[1]), ctypes.c_int32(int(weight_func_id)),
        _lib._dp(rb), ctypes.c_int32(len(rb)), ctypes.c_int64(first), ctypes.c_int64(last),
        _lib._dp(counts), extra_flags=_lib.cache_flags(c1, c2, PBCs))
    return np.array(_dist.allreduce_sum(counts))


def _process_one_weights(weights, npts_sample, correct_num_weights, weight_func_id, which):
    correct_shape = (npts_sample, correct_num_weights)
    converted = False
    if weights is None:
        weights = np.ones(correct_shape, dtype=np.float64)
    else:
        weights = np.atleast_1d(weights)
        weights = weights.astype("float64")
        if weights.ndim == 1:
            converted = True
            weights = weights.reshape((len(weights), 1))
        elif weights.ndim == 2:
            pass
        else:
            msg = ("\n You must either pass in a 1-D or 2-D array \n"
                   "for the input `weights%i`. Instead, an array of \n"
                   "dimension %i was received.")
            raise HalotoolsError(msg % (which, weights.ndim))
    npts_weights, num_weights = np.shape(weights)
    if np.shape(weights) != correct_shape:
        if converted is True:
            msg = ("\n You passed in a 1-D array for `weights%i` that \n"
                   "does not have the correct length. The number of \n"
                   "points in `sample%i` = %i, while the number of points \n"
                   "in your input 1-D `weights%i` array = %i")
            raise HalotoolsError(msg % (which, which, npts_sample, which, npts_weights))
        else:
            msg = ("\n You passed in a 2-D array for `weights%i` that \n"
                   "does not have a consistent shape with `sample%i`. \n"
                   "`sample%i` has length %i. The input value of `weight_func_id` = %i \n"
                   "For this value of `weight_func_id`, there should be %i weights \n"
                   "per point. The shape of your input `weights%i` is (%i, %i)\n")
            raise HalotoolsError(msg % (which, which, which, npts_sample, weight_func_id,
                                        correct_num_weights, which, npts_weights, num_weights))
    return weights


def _marked_npairs_process_weights(sample1, sample2, weights1, weights2, weight_func_id):
    """weights -> float64 (Npts, n_w) with n_w fixed by ``weight_func_id``; HalotoolsError on any
    shape mismatch (marked_npairs_3d.py:186-278)."""
    correct_num_weights = _func_signature_int_from_wfunc(weight_func_id)
    same = (weights2 is weights1) and (weights1 is not None) and (np.shape(sample1)[0] == np.shape(sample2)[0])
    weights1 = _process_one_weights(weights1, np.shape(sample1)[0], correct_num_weights, weight_func_id, 1)
    if same:
        return weights1, weights1
    weights2 = _process_one_weights(weights2, np.shape(sample2)[0], correct_num_weights, weight_func_id, 2)
    return weights1, weights2


def _func_signature_int_from_wfunc(weight_func_id):
    """Number of weights per point a weighting function reads (marked_npairs_3d.py:281-326)."""
    if type(weight_func_id) != int:
        msg = "\n weight_func_id parameter must be an integer ID of a weighting function."
        raise ValueError(msg)
    if weight_func_id in _NUM_WEIGHTS:
        return _NUM_WEIGHTS[weight_func_id]
    msg = ("The value ``weight_func_id`` = %i is not recognized")
    raise HalotoolsError(msg % weight_func_id)
