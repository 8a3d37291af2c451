"""ctypes binding of the C ABI in include/halotools_b200.h (libhalotools_b200.so).

The library is the product: there is NO CPU fallback.  Importing the package works
without a GPU (so argument validation can be tested anywhere) but every compute
call raises ``RuntimeError`` when the shared object is missing or no CUDA device
is visible.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HTB_LIB_PATH") or os.path.join(_HERE, "libhalotools_b200.so")   # override: kernel-variant sweeps

FLAG_DEVICE_INPUT = 1
FLAG_NO_CULL = 2
FLAG_GENERIC = 4
FLAG_NO_TMA = 8
FLAG_NO_SYM = 16
FLAG_UNIFORM_MASS = 32
FLAG_COLUMN_SUM = 64
FLAG_CACHE_SAMPLE1 = 128
FLAG_CACHE_SAMPLE2 = 256
FLAG_PARTITION_SUM = 512
FLAG_DEVICE_OUTPUT = 1024
FLAG_EARLY_EXIT = 2048
FLAG_PREPARE = 4096

EXPORTS = (
    "htb_last_error", "htb_abi_version", "htb_device_count", "htb_set_device", "htb_set_stream", "htb_set_shard",
    "htb_cache_begin", "htb_cache_end",
    "htb_npairs_3d_engine", "htb_npairs_xy_z_engine", "htb_npairs_s_mu_engine",
    "htb_marked_npairs_3d_engine", "htb_mean_delta_sigma_engine",
    "htb_marked_npairs_xy_z_engine", "htb_npairs_per_object_3d_engine", "htb_weighted_npairs_xy_engine",
    "htb_npairs_jackknife_3d_engine", "htb_npairs_jackknife_xy_z_engine", "htb_weighted_npairs_per_object_xy_engine",
    "htb_mesh_cell_ids", "htb_mesh_cell_id_indices", "htb_cell1_work", "htb_measure_fp64_rate",
    "htb_host_minmax", "htb_device_minmax", "htb_tp_estimator", "htb_get_stream", "htb_stream_synchronize", "htb_async_count_times", "htb_async_kernel_spans", "htb_async_kernel_stamps",
    "htb_return_xyz_formatted_array", "htb_apply_zspace_distortion", "htb_upload_f64",
)


class MeshGeom(ctypes.Structure):
    """htb_mesh_geom"""
    _fields_ = [("ndim", ctypes.c_int32), ("pbc", ctypes.c_int32),
                ("ndivs1", ctypes.c_int32 * 3), ("ndivs2", ctypes.c_int32 * 3),
                ("cover", ctypes.c_int32 * 3), ("reserved", ctypes.c_int32),
                ("period", ctypes.c_double * 3), ("cell1_size", ctypes.c_double * 3),
                ("cell2_size", ctypes.c_double * 3), ("search", ctypes.c_double * 3)]


class Stats(ctypes.Structure):
    """htb_stats"""
    _fields_ = [("pairs_evaluated", ctypes.c_double), ("pairs_reference", ctypes.c_double),
                ("ms_h2d", ctypes.c_float), ("ms_mesh", ctypes.c_float),
                ("ms_count", ctypes.c_float), ("ms_total", ctypes.c_float),
                ("kernel_launches", ctypes.c_int32), ("tiles", ctypes.c_int32),
                ("tiles_redone", ctypes.c_int32), ("refine1", ctypes.c_int32 * 3),
                ("refine2", ctypes.c_int32 * 3), ("path", ctypes.c_int32)]

    def as_dict(self):
        return {"pairs_evaluated": self.pairs_evaluated, "pairs_reference": self.pairs_reference,
                "ms_h2d": self.ms_h2d, "ms_mesh": self.ms_mesh, "ms_count": self.ms_count,
                "ms_total": self.ms_total, "kernel_launches": self.kernel_launches,
                "tiles": self.tiles, "tiles_redone": self.tiles_redone,
                "refine1": list(self.refine1), "refine2": list(self.refine2), "path": self.path}


_lib = None
last_stats = None        # Stats of the most recent engine call (dict), for benchmarks / tests
default_flags = 0        # OR-ed into every engine call (tests flip HTB_FLAG_GENERIC / NO_CULL here)
stream_flags = 0         # OR-ed into the engine calls a multi-stream statistic issues (HTB_FLAG_EARLY_EXIT)
collect_stats = True


def library_present():
    return os.path.exists(LIB_PATH)


def load():
    """Load the shared object (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "halotools_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.htb_last_error.restype = ctypes.c_char_p
        for name in EXPORTS:
            getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        _lib = lib
    return _lib


def require_gpu():
    lib = load()
    if lib.htb_device_count() < 1:
        raise RuntimeError("halotools_b200: no CUDA device is visible; the pair counters only run on a GPU "
                           "(there is no CPU fallback)")
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("halotools_b200: " + load().htb_last_error().decode("utf-8", "replace"))


def set_device(index):
    check(require_gpu().htb_set_device(int(index)))


def set_shard(rank, world):
    """This thread's engine calls count only rank's work-balanced share of the mesh1 cells (host state only)."""
    check(load().htb_set_shard(int(rank), int(world)))


class upload_cache(object):
    """Context manager: host samples uploaded by the engine calls inside stay on the device until exit
    (htb_cache_begin / htb_cache_end).  Re-entrant (only the outermost level frees)."""
    _depth = 0

    def __enter__(self):
        if library_present():
            if upload_cache._depth == 0:
                load().htb_cache_begin()
            upload_cache._depth += 1
            self._on = True
        else:
            self._on = False
        return self

    def __exit__(self, *exc):
        if self._on:
            upload_cache._depth -= 1
            if upload_cache._depth == 0:
                load().htb_cache_end()
        return False


def cache_flags(c1, c2, caller_owned):
    """Engine flags that let the samples of this call take part in the upload cache: only inside an
    ``upload_cache()`` block, and only for pointers into the caller's own arrays (``caller_owned``: the
    front-end made no shifted / converted temporary, whose address could be recycled)."""
    if upload_cache._depth == 0 or not caller_owned:
        return 0
    return (FLAG_CACHE_SAMPLE1 if c1.stable else 0) | (FLAG_CACHE_SAMPLE2 if c2.stable else 0)


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class Columns(object):
    """Column views of an (N, ndim) float64 sample: base pointers + common element stride.

    A C-contiguous (N, 3) array is passed as ONE block (x = base, y = base + 1, z = base + 2,
    stride 3) so the library uploads it with a single host->device copy.
    """

    def __init__(self, cols):
        self.device = all(getattr(c, "is_cuda", False) for c in cols)
        self.stable = False       # True: the pointers are views of the caller's own float64 array (no temporary copy)
        if self.device:
            # torch CUDA tensors (column views of an (N, ndim) float64 tensor): pass device pointers
            import torch
            n = int(cols[0].shape[0])
            if not all(c.dtype == torch.float64 and c.dim() == 1 and int(c.shape[0]) == n for c in cols):
                raise TypeError("device samples must be float64 CUDA tensors of shape (Npts, ndim)")
            strides = set(int(c.stride(0)) for c in cols) if n > 0 else {1}
            if len(strides) != 1:
                raise TypeError("device sample columns must share one stride")
            self.cols = cols
            self.n = n
            self.stride = list(strides)[0]
            self.ptrs = [ctypes.cast(ctypes.c_void_p(int(c.data_ptr())), ctypes.POINTER(ctypes.c_double)) for c in cols]
            return
        cols = [np.asarray(c) for c in cols]
        n = cols[0].shape[0]
        ok = all(c.dtype == np.float64 and c.ndim == 1 and c.shape[0] == n for c in cols)
        stride = None
        if ok and n > 0:
            strides = set(c.strides[0] for c in cols)
            ok = len(strides) == 1 and list(strides)[0] % 8 == 0 and list(strides)[0] > 0
            if ok:
                stride = list(strides)[0] // 8
        self.stable = bool(ok and n > 0)
        if not ok or n == 0:
            cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
            stride = 1
        self.cols = cols          # keep references alive for the duration of the call
        self.n = n
        self.stride = stride
        self.ptrs = [_dp(c) for c in cols]


def run_engine(func_name, *args, **kw):
    """Call an engine entry point, appending (flags, stats) and recording the stats.  ``out_device=True``: the
    output argument is a device pointer and the call is asynchronous (HTB_FLAG_DEVICE_OUTPUT; no stats)."""
    global last_stats
    lib = require_gpu()
    st = Stats()
    from . import distributed
    out_device = bool(kw.get("out_device"))
    flags = (default_flags | (FLAG_DEVICE_INPUT if kw.get("device") else 0) | int(kw.get("extra_flags", 0))
             | distributed.engine_flags() | (FLAG_DEVICE_OUTPUT if out_device else 0) | stream_flags)
    want_stats = collect_stats and not out_device
    rc = getattr(lib, func_name)(*args, ctypes.c_uint32(flags),
                                 ctypes.byref(st) if want_stats else None)
    check(rc)
    last_stats = st.as_dict() if want_stats else None
    return last_stats


def out_pointer(out, numpy_array, ctype):
    """ctypes pointer of an engine's output: the numpy array, or - asynchronous call - a torch CUDA tensor."""
    if out is None:
        return numpy_array.ctypes.data_as(ctypes.POINTER(ctype))
    if not (getattr(out, "is_cuda", False) and out.is_contiguous() and out.numel() == numpy_array.size
            and out.element_size() == 8):
        raise TypeError("device output must be a contiguous 8-byte CUDA tensor with %d elements" % numpy_array.size)
    return ctypes.cast(ctypes.c_void_p(int(out.data_ptr())), ctypes.POINTER(ctype))


_streams = {}


def engine_stream():
    """The CUDA stream this thread's engine calls are issued on, as a torch stream (plumbing: lets torch allocate the
    count tables and run the NCCL all-reduce in stream order with the kernels)."""
    import torch
    lib = require_gpu()
    ptr = ctypes.c_void_p()
    check(lib.htb_get_stream(ctypes.byref(ptr)))
    key = (torch.cuda.current_device(), ptr.value)
    if key not in _streams:
        _streams[key] = torch.cuda.ExternalStream(ptr.value or 0, device=torch.cuda.current_device())
    return _streams[key]


def async_count_times():
    """CUDA-event durations (ms) of the counting kernels of the asynchronous calls since the last query."""
    lib = require_gpu()
    buf = (ctypes.c_float * 16)()
    n = ctypes.c_int32(0)
    check(lib.htb_async_count_times(buf, ctypes.c_int32(16), ctypes.byref(n)))
    return [float(buf[i]) for i in range(n.value)]


def async_kernel_spans():
    """Device-side durations (ms; first warp in -> last warp out) of the counting kernels of the asynchronous calls since
    the last ``async_count_times`` query (which resets the ring: ask for the spans first)."""
    lib = require_gpu()
    buf = (ctypes.c_float * 16)()
    n = ctypes.c_int32(0)
    check(lib.htb_async_kernel_spans(buf, ctypes.c_int32(16), ctypes.byref(n)))
    return [float(buf[i]) for i in range(n.value)]


def async_kernel_stamps():
    """(first warp in, last warp out) in ns of the device's globaltimer for the same launches (not reset either)."""
    lib = require_gpu()
    buf = (ctypes.c_uint64 * 32)()
    n = ctypes.c_int32(0)
    check(lib.htb_async_kernel_stamps(buf, ctypes.c_int32(16), ctypes.byref(n)))
    return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value)]


class use_stream(object):
    """Context manager: this thread's engine calls inside are issued on ``stream`` (a torch.cuda.Stream); the previous
    stream is restored on exit."""

    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        lib = require_gpu()
        prev = ctypes.c_void_p()
        check(lib.htb_get_stream(ctypes.byref(prev)))
        self.prev = prev.value
        check(lib.htb_set_stream(ctypes.c_void_p(self.stream.cuda_stream)))
        return self.stream

    def __exit__(self, *exc):
        check(load().htb_set_stream(ctypes.c_void_p(self.prev)))
        return False


def uploadable(a):
    """A host sample the library can bring to the device as it is: C-contiguous float64 (Npts, ndim)."""
    return (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == 2 and a.flags.c_contiguous
            and a.shape[0] > 0)


def upload_rows(host, dev):
    """Enqueue the copy of a C-contiguous float64 host array into the CUDA tensor ``dev`` (same number of elements) on
    the engine's stream; the host array must stay alive until the stream is synchronised."""
    assert host.size == dev.numel() and dev.is_contiguous()
    check(require_gpu().htb_upload_f64(_dp(host), ctypes.c_int64(host.size),
                                       ctypes.cast(ctypes.c_void_p(int(dev.data_ptr())), ctypes.POINTER(ctypes.c_double))))


def stream_synchronize():
    check(require_gpu().htb_stream_synchronize())


def measure_fp64_rate():
    lib = require_gpu()
    rate = ctypes.c_double(0.0)
    clk = ctypes.c_double(0.0)
    check(lib.htb_measure_fp64_rate(ctypes.byref(rate), ctypes.byref(clk)))
    return rate.value, clk.value
