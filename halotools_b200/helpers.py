"""Argument helpers shared by the front-ends.

Same behaviour (return values, exception types and message text the reference's tests
assert on) as /root/reference/halotools/mock_observables/mock_observables_helpers.py:25-200
and /root/reference/halotools/utils/array_utils.py:14,135,189.
"""
import multiprocessing
from warnings import warn

import numpy as np

from .custom_exceptions import HalotoolsError

num_available_cores = multiprocessing.cpu_count()

__all__ = ("custom_len", "array_is_monotonic", "unsorting_indices", "get_num_threads", "get_period",
           "enforce_sample_respects_pbcs", "enforce_sample_has_correct_shape",
           "get_separation_bins_array", "get_line_of_sight_bins_array")


def custom_len(x):
    """len() that answers 1 for scalars (array_utils.py:14)."""
    try:
        return len(x)
    except TypeError:
        return 1


def array_is_monotonic(array, strict=False):
    """+1 increasing, -1 decreasing, 0 neither (array_utils.py:135)."""
    if custom_len(array) < 3:
        raise HalotoolsError("Input array to the array_is_monotonic method has less then 3 elements")
    d = np.diff(array)
    if strict is True:
        up, down = np.all(d > 0), np.all(d < 0)
    else:
        up, down = np.all(d >= 0), np.all(d <= 0)
    return 1 if up else (-1 if down else 0)


def unsorting_indices(sorting_indices):
    """Inverse permutation of an argsort (array_utils.py:189)."""
    out = np.zeros_like(sorting_indices, dtype=int)
    out[sorting_indices] = np.arange(len(sorting_indices)).astype(int)
    return out


def check_num_threads_arg(num_threads):
    """The check every pair-counter front-end makes (npairs_3d.py:157-162): ``num_threads`` is
    validated exactly like the reference and then IGNORED - the GPU engine has no thread knob."""
    if num_threads != 1 or not isinstance(num_threads, int):
        if isinstance(num_threads, str) and num_threads == "max":
            num_threads = multiprocessing.cpu_count()
        if not isinstance(num_threads, int):
            raise ValueError("Input ``num_threads`` argument must be an integer or the string 'max'")
    return num_threads


def get_num_threads(input_num_threads, enforce_max_cores=False):
    """mock_observables_helpers.py:74."""
    if isinstance(input_num_threads, str) and input_num_threads == "max":
        num_threads = num_available_cores
    else:
        try:
            num_threads = int(input_num_threads)
            assert num_threads == input_num_threads
        except Exception:
            raise ValueError("Input ``num_threads`` must be an integer")
    if num_threads > num_available_cores:
        msg = "Input ``num_threads`` = {0} exceeds the ``num_available_cores`` = {1}.\n"
        warn(msg.format(num_threads, num_available_cores))
        if enforce_max_cores is True:
            warn("Since ``enforce_max_cores`` is True,\nsetting ``num_threads`` to ``num_available_cores``.")
            num_threads = num_available_cores
    return num_threads


def get_period(period):
    """(period, PBCs): None -> (None, False); scalar / 3-sequence -> (float[3], True)
    (mock_observables_helpers.py:107)."""
    if period is None:
        return period, False
    period = np.atleast_1d(period).astype(float)
    if len(period) == 1:
        period = np.array([period[0]] * 3).astype(float)
    if not (np.all(period < np.inf) and np.all(period > 0) and len(period) == 3):
        raise ValueError("Input ``period`` must be either a scalar or a 3-element sequence.\n"
                         "All values must bounded positive numbers.\n")
    return period, True


def _column_extrema(x, y, z):
    """(min, max) per coordinate.  Large float64 samples whose three coordinates are adjacent columns of one
    matrix are scanned once by the library's threaded host helper; anything else goes through numpy."""
    if all(getattr(a, "is_cuda", False) for a in (x, y, z)):
        # device-resident columns (torch CUDA tensors): the extrema are found where the data lives - in ONE pass
        # when the three columns are views of one (N, 3) tensor (the usual case), with one small copy to the host
        import torch
        if x.numel() == 0:
            return [(0.0, 0.0)] * 3
        base = getattr(x, "_base", None)
        if (base is not None and base.dim() == 2 and base.shape[1] == 3 and y._base is base and z._base is base
                and x.data_ptr() == base.data_ptr() and y.data_ptr() == base[:, 1].data_ptr()
                and z.data_ptr() == base[:, 2].data_ptr()):
            if base.dtype == torch.float64 and base.is_contiguous():
                # one streaming pass of the library's own kernel (HBM bound; torch's strided reductions are not)
                import ctypes
                from . import _lib
                lo = (ctypes.c_double * 3)()
                hi = (ctypes.c_double * 3)()
                _lib.check(_lib.require_gpu().htb_device_minmax(ctypes.c_void_p(int(base.data_ptr())), ctypes.c_int64(int(base.shape[0])),
                                                                ctypes.c_int64(3), ctypes.c_int32(3), lo, hi))
                return [(lo[k], hi[k]) for k in range(3)]
            lo, hi = torch.aminmax(base, dim=0)
            ext = torch.stack([lo, hi]).cpu().numpy()
            return [(float(ext[0, k]), float(ext[1, k])) for k in range(3)]
        return [(float(a.min()), float(a.max())) for a in (x, y, z)]
    try:
        n = len(x)
        if (n >= 1000000 and all(isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == 1 for a in (x, y, z))
                and len(y) == n and len(z) == n and x.strides == y.strides == z.strides and x.strides[0] % 8 == 0
                and x.strides[0] >= 24
                and y.ctypes.data - x.ctypes.data == 8 and z.ctypes.data - y.ctypes.data == 8):
            import ctypes
            from . import _lib
            lib = _lib.load()
            lo = (ctypes.c_double * 3)()
            hi = (ctypes.c_double * 3)()
            rc = lib.htb_host_minmax(ctypes.c_void_p(x.ctypes.data), ctypes.c_int64(n), ctypes.c_int64(x.strides[0] // 8),
                                     ctypes.c_int32(3), lo, hi)
            if rc == 0:
                return [(lo[k], hi[k]) for k in range(3)]
    except Exception:
        pass
    return [(np.min(a), np.max(a)) if len(a) else (0.0, 0.0) for a in (x, y, z)]


def enforce_sample_respects_pbcs(x, y, z, period):
    """0 <= coordinate <= period in every dimension (mock_observables_helpers.py:25).  The reference tests
    ``np.all(x >= 0)`` / ``np.all(x <= period)``; the same decisions are taken here from the column extrema
    (a NaN fails both tests, as it does in the reference)."""
    ext = _column_extrema(x, y, z)
    if not all(lo >= 0 for lo, _ in ext):
        msg = ("You set periodic boundary conditions to be True by passing in \n"
               "period = (%.2f, %.2f, %.2f), but your input data has negative values,\n"
               "indicating that you forgot to apply periodic boundary conditions.\n")
        raise ValueError(msg % (period[0], period[1], period[2]))
    for (_, hi), name, p in zip(ext, ("x", "y", "z"), (period[0], period[1], period[2])):
        if not hi <= p:
            msg = ("You set %speriod = %.2f but there are values in the %s-dimension \n"
                   "of the input data that exceed this value")
            raise ValueError(msg % (name, p, name))


def enforce_sample_has_correct_shape(sample, ndim=3):
    """(Npts, ndim) or TypeError (mock_observables_helpers.py:135)."""
    if not getattr(sample, "is_cuda", False):
        sample = np.atleast_1d(sample)
    shape = tuple(sample.shape)
    if not (len(shape) == 2 and shape[1] == ndim):
        msg = ("Input sample of points must be a Numpy ndarray of shape (Npts, {0}).\n"
               "To convert a sequence of 1d arrays x, y, z into correct shape expected \n"
               "throughout the `mock_observables` package:\n\n"
               ">>> sample = np.vstack([x, y, z]).T ".format(ndim))
        raise TypeError(msg)
    return sample


def _bins_ok(bins, positive):
    try:
        assert bins.ndim == 1
        assert len(bins) > 1
        if len(bins) > 2:
            assert array_is_monotonic(bins, strict=True) == 1
        if positive:
            assert np.all(bins > 0)
    except AssertionError:
        return False
    return True


def get_separation_bins_array(separation_bins):
    """Strictly positive, strictly increasing 1-d bins or TypeError (mock_observables_helpers.py:153)."""
    separation_bins = np.atleast_1d(separation_bins)
    if not _bins_ok(separation_bins, True):
        raise TypeError("\n Input separation bins must be a monotonically increasing \n"
                        "1-D array with at least two entries, all of which must be strictly positive.\n")
    return separation_bins


def get_line_of_sight_bins_array(pi_bins):
    """Like get_separation_bins_array but zero is allowed (mock_observables_helpers.py:180)."""
    pi_bins = np.atleast_1d(pi_bins)
    if not _bins_ok(pi_bins, False):
        raise TypeError("\n Input separation bins must be a monotonically increasing \n"
                        "1-D array with at least two entries.\n")
    return pi_bins
