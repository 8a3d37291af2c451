"""K3 on the device: the pair counts of one statistic stay in HBM from the counting kernels to the estimator.

``tpcf`` / ``rp_pi_tpcf`` / ``wp`` make up to six engine calls (D1D1, D1D2, D2D2, D1R, D2R, RR;
/root/reference/halotools/mock_observables/two_point_clustering/tpcf.py:76-113,164-205, rp_pi_tpcf.py:296-467) and then
combine np.diff's of the counts (tpcf_estimators.py:14-119, wp.py:219-221).  Here every count is ENQUEUED on the
engine's CUDA stream (HTB_FLAG_DEVICE_OUTPUT) and writes its cumulative table into one device buffer; the ranks' partial
tables are summed by ONE all-reduce issued on the same stream (NCCL through torch.distributed: plumbing), the estimator
kernel (htb_tp_estimator) reads the summed tables, and the host waits ONCE, for the D2H copy of xi and of the
zero-division flags.  Host samples are brought to the device ONCE per statistic (each rank 1/world of the rows + an
all-gather, ``distributed.to_device``), and the engine calls then run on SEPARATE streams, the largest count first: every
count kernel is a persistent grid whose last, partly filled round of tiles leaves SM slots idle (4.2 tiles per warp on one
rank of 8 take the time of 5); with HTB_FLAG_EARLY_EXIT the surplus blocks retire at once and idle warps do not linger, so
the set-up and count kernels of the smaller counts run in those slots.  Orders and set-up / count splits that were
measured and lost (issue order, all set-ups first, the smaller counts' set-ups first) are listed in DESIGN.md section 5.
The numpy estimators of ``tpcf_estimators.py`` remain for the jackknife statistics (rows per
sub-volume) and as the host restatement the CPU tests of the driver logic run.
"""
import ctypes
import os

import numpy as np

from .. import _lib
from .. import distributed as _dist
from .tpcf_estimators import _ZERO_MSG, _list_estimators

__all__ = ("DeviceStatistic", "available")

MAX_TABLES = 6
TIMELINE = None          # debugging: a list makes every enqueue append (what, stream index, torch event recorded behind it)
SIDE_STREAMS = 3
_side = {}


def _side_streams(torch):
    dev = torch.cuda.current_device()
    if dev not in _side:
        _side[dev] = [torch.cuda.Stream(device=dev) for _ in range(SIDE_STREAMS)]
    return _side[dev]


def available(counter):
    """True when ``counter`` (the module-level pair counter of a statistic) is the GPU front-end with an ``enqueue``
    entry - the CPU tests of the host layer replace it by a checker function, which has none."""
    return getattr(counter, "enqueue", None) is not None and _lib.library_present()


def _ptr(t, ctype):
    if t is None:
        return None
    return ctypes.cast(ctypes.c_void_p(int(t.data_ptr())), ctypes.POINTER(ctype))


class DeviceStatistic(object):
    """Cumulative count tables (int64, shape ``shape``) of one statistic in one device buffer, plus the xi rows the
    estimator kernel fills."""

    def __init__(self, shape, max_xi=3):
        import torch
        self.torch = torch
        _lib.require_gpu()
        self.stream = _lib.engine_stream()
        self.shape = tuple(int(v) for v in shape)
        self.ncum = int(np.prod(self.shape))
        self.n0 = self.shape[0]
        self.n1 = self.shape[1] if len(self.shape) > 1 else 1
        self.nout = (self.n0 - 1) * max(self.n1 - 1, 1)
        with torch.cuda.stream(self.stream):
            self.tables = torch.zeros((MAX_TABLES, self.ncum), dtype=torch.int64, device="cuda")
            # xi rows, then one row whose first entries are the zero-division flags (int32 view)
            self.xi = torch.zeros((max_xi + 1, max(self.nout, 2)), dtype=torch.float64, device="cuda")
            # uploads that follow on this stream are awaited per sample, not wholesale (launch): the tables' own event
            self.tables_ready = torch.cuda.Event()
            self.tables_ready.record(self.stream)
        self.uploaded = {}      # data_ptr of a sample uploaded by inputs() -> (order of arrival, event behind its copy)
        self.flags = self.xi[max_xi].view(torch.int32)
        self.used = 0
        self.nxi = 0
        self.keep = []          # host arrays / analytic tables the enqueued work still reads
        self.reduced = False
        self.estimators = []
        self.multi = False      # engine calls on side streams (all samples device resident)
        self.forked = []
        self.deferred = []

    def inputs(self, periodic, *samples):
        """Bring the statistic's samples to the device once (None entries and repeated objects keep their identity).
        Host samples that cannot be copied as they are (other dtypes, strided views), and the samples of a NON-periodic
        call (the front-ends shift those into an enclosing box on the host, mesh_helpers.py:17-64), stay on the host: the
        engines then upload them themselves (upload cache) and the calls share one stream."""
        if not periodic:
            return tuple(samples)
        distinct = {}
        for a in samples:
            if a is not None:
                distinct[id(a)] = a
        arrays = list(distinct.values())
        on_device = [getattr(a, "is_cuda", False) for a in arrays]
        ok = all(c or _lib.uploadable(a) for a, c in zip(arrays, on_device))
        if ok and arrays:
            with self.torch.cuda.stream(self.stream):
                for a, c in zip(arrays, on_device):
                    if not c:
                        dev = distinct[id(a)] = _dist.to_device(a)
                        self.keep.append(a)
                        ev = self.torch.cuda.Event()
                        ev.record(self.stream)
                        self.uploaded[int(dev.data_ptr())] = (len(self.uploaded), ev)
            self.multi = not os.environ.get("HTB_ONE_STREAM")
        return tuple(None if a is None else distinct[id(a)] for a in samples)

    def count(self, enqueue, *args, **kwargs):
        """Run ``enqueue(out, *args, **kwargs)`` into the next free table; returns the table (a device int64 row)."""
        if self.used >= MAX_TABLES:
            raise RuntimeError("DeviceStatistic: more than %d count tables" % MAX_TABLES)
        out = self.tables[self.used]
        k = self.used
        self.used += 1
        if not self.multi:
            with self.torch.cuda.stream(self.stream):
                self.keep.append(enqueue(out, *args, **kwargs))
            return out
        # multi-stream mode: nothing is enqueued yet - launch() issues the counts largest first
        a, b = args[0], args[1]
        self.deferred.append((int(a.shape[0]) * int(b.shape[0]), k, enqueue, out, args, kwargs))
        return out

    def _enqueue(self, side, flag, enqueue, out, args, kwargs):
        saved = _lib.stream_flags
        _lib.stream_flags = saved | flag
        try:
            with self.torch.cuda.stream(side), _lib.use_stream(side):
                self.keep.append(enqueue(out, *args, **kwargs))
        finally:
            _lib.stream_flags = saved
        if TIMELINE is not None:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record(side)
            TIMELINE.append(("prepare" if flag == _lib.FLAG_PREPARE else "count", _side_streams(self.torch).index(side), ev))

    def launch(self):
        """Multi-stream mode: every count is one engine call (set-up + persistent count kernel) on its own stream, the
        LARGEST first (HTB_FLAG_EARLY_EXIT: the surplus blocks of a launch retire at once and idle warps do not wait, so the
        set-up kernels and count kernels of the smaller counts run beside the big one - in the SM slots its last,
        partly filled round of tiles would have left idle - instead of before or after it)."""
        deferred, self.deferred = self.deferred, []

        def arrival(d):
            # order in which the samples of a count arrive on the device (0: resident, or the first upload)
            return max([self.uploaded[int(t.data_ptr())][0] for t in d[4][:2]
                        if getattr(t, "is_cuda", False) and int(t.data_ptr()) in self.uploaded] or [0])

        if not os.environ.get("HTB_ISSUE_ORDER"):
            # largest first; while samples are still crossing PCIe, the counts whose samples arrive first go first (tpcf
            # from host arrays: DD runs while the randoms are being copied)
            deferred.sort(key=lambda d: (arrival(d), -d[0]) if self.uploaded else -d[0])
        sides = []
        for i in range(len(deferred)):
            side = _side_streams(self.torch)[i % SIDE_STREAMS]
            if side not in self.forked:
                if self.uploaded:
                    side.wait_event(self.tables_ready)
                else:
                    side.wait_stream(self.stream)      # the samples and the zeroed tables were produced on the main stream
                self.forked.append(side)
            sides.append(side)
        for i, (_size, _k, enqueue, out, args, kwargs) in enumerate(deferred):
            if self.uploaded:
                # wait for THIS count's samples only (a sample that was not uploaded here was resident before the call)
                for t in args[:2]:
                    if getattr(t, "is_cuda", False) and int(t.data_ptr()) in self.uploaded:
                        sides[i].wait_event(self.uploaded[int(t.data_ptr())][1])
            self._enqueue(sides[i], _lib.FLAG_EARLY_EXIT, enqueue, out, args, kwargs)

    def analytic(self, array):
        """Differential float counts computed on the host (analytic randoms) -> device row."""
        a = np.ascontiguousarray(array, dtype=np.float64).reshape(-1)
        assert a.size == self.nout
        with self.torch.cuda.stream(self.stream):
            t = self.torch.from_numpy(a).to("cuda", non_blocking=False)
        self.keep.append((a, t))
        return t

    def reduce(self):
        """Sum the ranks' partial tables: one all-reduce on the device buffer, on the engine's stream."""
        if self.reduced:
            return
        self.reduced = True
        self.launch()
        for side in self.forked:
            self.stream.wait_stream(side)
        rank, world = _dist._rank_world()
        if world > 1 and self.used > 0:
            import torch.distributed as dist
            with self.torch.cuda.stream(self.stream):
                dist.all_reduce(self.tables[:self.used], op=dist.ReduceOp.SUM, group=_dist._state["group"])

    def estimator(self, DD, D1R, D2R, RR, ND1, ND2, NR1, NR2, estimator, cross=False, wp_pi_max=0.0):
        """Enqueue _TP_estimator / _TP_estimator_crossx (tpcf_estimators.py:14-119) over device tables; operands are
        int64 tables from ``count`` or float rows from ``analytic`` (None: not needed by this estimator).  Returns the
        index of the xi row."""
        self.reduce()
        names = _list_estimators()
        if estimator not in names:
            raise ValueError("unsupported estimator!")
        if cross and estimator not in ("Natural", "Hamilton", "Landy-Szalay"):
            raise ValueError("{0} estimator is not supported for cross-correlations".format(estimator))
        code = names.index(estimator)
        # the normalisation factors exactly as the reference forms them (numpy integer products, one true division)
        ND1, ND2, NR1, NR2 = (np.atleast_1d(v) for v in (ND1, ND2, NR1, NR2))
        with np.errstate(divide="ignore", invalid="ignore"):
            if estimator == "Davis-Peebles":
                f1 = (1.0 / (ND1 * ND2 / (ND1 * NR2)))[0]
                f2 = 0.0
            else:
                f1 = (1.0 / (ND1 * ND2 / (NR1 * NR2)))[0]
                f2 = (1.0 / (ND1 * NR2 / (NR1 * NR2)))[0]

        def split(op):
            if op is None:
                return None, None
            return (op, None) if op.dtype == self.torch.int64 else (None, op)

        dd_c, dd_f = split(DD)
        if dd_c is None:
            raise TypeError("DD must be a counted table")
        d1_c, d1_f = split(D1R)
        d2_c, d2_f = split(D2R)
        rr_c, rr_f = split(RR)
        row = self.nxi
        self.nxi += 1
        i64, f64 = ctypes.c_int64, ctypes.c_double
        lib = _lib.require_gpu()
        with self.torch.cuda.stream(self.stream):
            _lib.check(lib.htb_tp_estimator(
                ctypes.c_int32(self.n0), ctypes.c_int32(self.n1), ctypes.c_int32(code), ctypes.c_int32(1 if cross else 0),
                _ptr(dd_c, i64), _ptr(d1_c, i64), _ptr(d2_c, i64), _ptr(rr_c, i64),
                _ptr(d1_f, f64), _ptr(d2_f, f64), _ptr(rr_f, f64),
                ctypes.c_double(float(f1)), ctypes.c_double(float(f2)), ctypes.c_double(float(wp_pi_max)),
                _ptr(self.xi[row], f64), _ptr(self.flags, ctypes.c_int32)))
        self.estimators.append(estimator)
        return row

    def fetch(self, out_shape=None):
        """The ONE host synchronisation of the statistic: D2H of the xi rows + flags; raises the reference's
        zero-division ValueError (tpcf_estimators.py:165-183).  Returns the list of xi arrays."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            host = self.xi.cpu()            # D2H on the engine's stream + wait
        self.keep = []
        flags = int(host[-1].view(torch.int32)[0])
        est = self.estimators[0] if self.estimators else "Natural"
        if flags & 1:
            raise ValueError(_ZERO_MSG.format("RR", est))
        if flags & 2:
            raise ValueError(_ZERO_MSG.format("DR", est))
        rows = host.numpy()
        shape = out_shape if out_shape is not None else ((self.n0 - 1,) if self.n1 == 1 else (self.n0 - 1, self.n1 - 1))
        return [np.array(rows[i, :self.nout]).reshape(shape) for i in range(self.nxi)]

    def cumulative(self):
        """Host copies of the (reduced) cumulative tables - for callers that want the raw counts (tests, bench)."""
        self.reduce()
        with self.torch.cuda.stream(self.stream):
            return self.tables[:self.used].cpu().numpy().reshape((self.used,) + self.shape)


def combine(stat, same, do_auto, do_cross, D1D1, D1D2, D2D2, D1R, D2R, RR, N1, N2, NR, estimator, wp_pi_max=0.0,
            out_shape=None):
    """``_driver.combine`` on the device: the same case analysis (tpcf.py:428-498), estimator kernels instead of numpy,
    one synchronisation for all returned arrays."""
    def auto(dd, dr, n):
        return stat.estimator(dd, dr, None, RR, n, n, NR, NR, estimator, wp_pi_max=wp_pi_max)

    def cross():
        return stat.estimator(D1D2, D1R, D2R, RR, N1, N2, NR, NR, estimator, cross=True, wp_pi_max=wp_pi_max)

    if same:
        auto(D1D1, D1R, N1)
        return stat.fetch(out_shape)[0]
    if (do_auto is True) & (do_cross is True):
        auto(D1D1, D1R, N1)
        cross()
        auto(D2D2, D2R, N2)
        xi = stat.fetch(out_shape)
        return xi[0], xi[1], xi[2]
    elif do_cross is True:
        cross()
        return stat.fetch(out_shape)[0]
    elif do_auto is True:
        auto(D1D1, D1R, N1)
        auto(D2D2, D2R, N2)
        xi = stat.fetch(out_shape)
        return xi[0], xi[1]
