"""One argument processor for every pair-counter front-end.

The reference repeats the same block in each front-end (pair_counters/npairs_3d.py:153-213, npairs_xy_z.py:166-237,
npairs_projected.py:160-227, surface_density/weighted_npairs_xy.py:153-217, ...): thread argument, column views, bin
checks, period / enclosing box, default cell sizes.  Here it exists once, parameterised by the number of mesh
dimensions, the list of bin arrays and the search length each one implies; the checks, their order and the error
strings are the reference's (its tests assert on the messages).
"""
import numpy as np

from .. import _lib
from ..helpers import array_is_monotonic, check_num_threads_arg, custom_len
from .mesh_helpers import _enclose_in_box, _enclose_in_square

__all__ = ("process_counter_args", "sample_columns")

_BINS_MSG = "Input ``%s`` must be a monotonically increasing 1D array with at least two entries"
_PERIOD_MSG = "Input ``period`` must be a bounded positive number in all dimensions"


def _checked_bins(values, name):
    bins = np.atleast_1d(values).astype('f8')
    ok = bins.ndim == 1 and len(bins) > 1
    if ok and len(bins) > 2:
        ok = array_is_monotonic(bins, strict=True) == 1
    if not ok:
        raise ValueError(_BINS_MSG % name)
    return bins


def _triple(value, default, ndim):
    if value is None:
        return list(default)
    if custom_len(value) == 1:
        return [value] * ndim
    return value


def process_counter_args(ndim, sample1, sample2, bins, reach, period, num_threads,
                         approx_cell1_size, approx_cell2_size, cell_default=None, after_period=None):
    """Returns (cols1, cols2, checked_bins, period, num_threads, PBCs, approx_cell1_size, approx_cell2_size).

    ``bins``: [(values, name), ...] checked in order; ``reach(checked)`` -> the ``ndim`` search lengths (3 x them is the
    minimum side of the non-periodic enclosing box, and - unless ``cell_default(checked)`` says otherwise - the default
    cell size); ``after_period(checked, period)``: extra checks that need the period (npairs_projected's pi_max).
    cols1 / cols2 are lists of ``ndim`` column views; they are the same list when sample2 is sample1."""
    num_threads = check_num_threads_arg(num_threads)
    cols1 = [sample1[:, d] for d in range(ndim)]
    cols2 = cols1 if sample2 is sample1 else [sample2[:, d] for d in range(ndim)]
    checked = [_checked_bins(v, name) for v, name in bins]
    search = list(reach(checked))
    if period is None:
        if getattr(sample1, "is_cuda", False) or getattr(sample2, "is_cuda", False):
            raise ValueError("device-resident samples need an explicit ``period``")
        PBCs = False
        enclose = _enclose_in_box if ndim == 3 else _enclose_in_square
        out = enclose(*(cols1 + cols2), min_size=[3.0 * s for s in search])
        shifted1, shifted2, period = list(out[:ndim]), list(out[ndim:2 * ndim]), out[2 * ndim]
        cols2 = shifted1 if cols2 is cols1 else shifted2
        cols1 = shifted1
    else:
        PBCs = True
        period = np.atleast_1d(period).astype(float)
        if len(period) == 1:
            period = np.array([period[0]] * ndim)
        if not (np.all(period < np.inf) and np.all(period > 0)):
            raise ValueError(_PERIOD_MSG)
    if after_period is not None:
        after_period(checked, period)
    default = search if cell_default is None else cell_default(checked)
    return (cols1, cols2, checked, period, num_threads, PBCs,
            _triple(approx_cell1_size, default, ndim), _triple(approx_cell2_size, default, ndim))


def sample_columns(cols1, cols2, host_only=None):
    """``_lib.Columns`` of both samples (one object when they are the same arrays).  Both samples must live on the
    same side of PCIe; ``host_only``: name of a counter whose weights / tags are host arrays."""
    c1 = _lib.Columns(list(cols1))
    same = cols2 is cols1 or all(a is b for a, b in zip(cols1, cols2))
    c2 = c1 if same else _lib.Columns(list(cols2))
    if c1.device != c2.device:
        raise TypeError("sample1 and sample2 must both be host arrays or both be CUDA tensors")
    if c1.device and host_only:
        raise TypeError("%s takes host (numpy) samples: its per-point weights are host arrays" % host_only)
    return c1, c2
