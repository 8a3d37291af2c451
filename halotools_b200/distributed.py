"""Multi-GPU plumbing: one process per GPU (torchrun), ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests).

The pair-counting path shards the way the reference does (contiguous ranges of mesh1 cells,
/root/reference/halotools/mock_observables/pair_counters/mesh_helpers.py:183-221, summed at the
end, npairs_3d.py:145): every rank holds all of sample2, counts the pairs of ITS range of
reference mesh1 cells, and the tiny count vectors are combined with ONE all-reduce (int64 sums are
exact, so results stay bit-identical to a single-GPU run).

Sharding is opt-in: call ``enable()`` after ``torch.distributed.init_process_group``; without it
every call counts all cells (N independent replicas).
"""
import numpy as np

_state = {"enabled": False, "group": None, "balanced": False, "local": 0, "cells": None}


def enable(group=None, balanced=None):
    """Shard subsequent pair-counter calls over the ranks of ``group`` (default: world).

    ``balanced`` (default: True on NCCL, i.e. when the GPUs are there): the engine cuts the mesh1
    cell range on the device so that every rank gets the same predicted work (htb_set_shard); the
    front-ends then pass the full cell range.  Otherwise the reference's equal-cell-count rule
    (``cell1_range``) is applied on the host."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state["enabled"] = True
    _state["group"] = group
    if balanced is None:
        balanced = dist.get_backend(group) == "nccl"
    _state["balanced"] = bool(balanced)
    if _state["balanced"]:
        from . import _lib
        _lib.set_shard(dist.get_rank(group), dist.get_world_size(group))


def disable():
    if _state["balanced"]:
        from . import _lib
        _lib.set_shard(0, 1)
    _state["enabled"] = False
    _state["group"] = None
    _state["balanced"] = False


def is_enabled():
    return _state["enabled"]


def _rank_world():
    if not _state["enabled"]:
        return 0, 1
    import torch.distributed as dist
    return dist.get_rank(_state["group"]), dist.get_world_size(_state["group"])


def device_collective():
    """True when the ranks' results can be summed where they are: sharding enabled over NCCL (the GPUs are there).  The
    front-ends then leave their outputs on the device (HTB_FLAG_DEVICE_OUTPUT) and all-reduce the device buffer on the
    engine's stream - no D2H -> numpy -> H2D hop (the analogue of the reference's sum over its worker pool,
    pair_counters/npairs_3d.py:145)."""
    if not _state["enabled"] or _state["local"] > 0:
        return False
    import torch.distributed as dist
    return dist.get_world_size(_state["group"]) > 1 and dist.get_backend(_state["group"]) == "nccl"


def allreduce_device(tensor):
    """Sum a CUDA tensor over the ranks in place, on the current (the engine's) stream."""
    import torch.distributed as dist
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=_state["group"])
    return tensor


def to_device(sample):
    """A C-contiguous float64 host sample (N, d) as a CUDA tensor, copied on the engine's stream (asynchronous: the
    caller keeps ``sample`` alive until it synchronises).  With sharding over NCCL every rank uploads only ITS 1/world of
    the rows across PCIe and one all-gather over NVLink / NVSwitch completes every rank's copy (every rank holds the
    same host array: the reference's workers all see the whole sample too, npairs_3d.py:139-146)."""
    import torch
    from . import _lib
    n, d = sample.shape
    if not device_collective():
        dev = torch.empty((n, d), dtype=torch.float64, device="cuda")
        _lib.upload_rows(sample, dev)
        return dev
    import torch.distributed as dist
    rank, world = _rank_world()
    per = -(-n // world)
    full = torch.empty((per * world, d), dtype=torch.float64, device="cuda")
    a, b = min(n, rank * per), min(n, (rank + 1) * per)
    part = full[rank * per:(rank + 1) * per]
    if b > a:
        _lib.upload_rows(sample[a:b], part[:b - a])
    dist.all_gather_into_tensor(full, part, group=_state["group"])        # in place: part is this rank's slot of full
    return full[:n]


def split_cells(ncells, world, work=None):
    """Contiguous (first, last) mesh1-cell ranges for ``world`` ranks.

    Without ``work`` this is the reference's np.array_split rule; with a per-cell predicted work
    vector (htb_cell1_work) the cut points equalise cumulative work instead of cell counts."""
    if world <= 1:
        return [(0, ncells)]
    if work is None:
        parts = np.array_split(np.arange(ncells), world)
        out, pos = [], 0
        for p in parts:
            out.append((pos, pos + len(p)))
            pos += len(p)
        return out
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(work, dtype=np.float64))])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(ncells)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, ncells))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


class cell_range(object):
    """Context manager: the pair counters inside count only the reference mesh1 cells [first, last) - the
    ``cell1_tuple`` every reference engine takes (npairs_3d_engine.pyx:17,103), which the reference front-ends
    never expose.  Used to compare full-size counts with the reference on a sample of its cells."""

    def __init__(self, first, last):
        self.range = (int(first), int(last))

    def __enter__(self):
        self.saved = _state["cells"]
        _state["cells"] = self.range
        return self

    def __exit__(self, *exc):
        _state["cells"] = self.saved
        return False


def engine_flags():
    """HTB_FLAG_PARTITION_SUM when the cell range an engine call receives is this rank's part of a partition whose
    counts are summed over the ranks (the symmetric auto-correlation shortcut is then allowed on a partial range); an
    explicit ``cell_range`` block asks for the reference's own per-range counts instead."""
    if _state["enabled"] and _state["cells"] is None and _rank_world()[1] > 1:
        return 512
    return 0


def cell1_range(ncells, work=None):
    """This rank's (first_cell1, last_cell1); the full range when the engine does the (balanced) cut."""
    if _state["cells"] is not None:
        return max(0, _state["cells"][0]), min(ncells, _state["cells"][1])
    if _state["balanced"]:
        return 0, ncells
    rank, world = _rank_world()
    return split_cells(ncells, world, work)[rank]


class local_counts(object):
    """Context manager for statistics that make several engine calls (tpcf: DD, DR, RR): inside it the pair counters
    return this rank's partial counts (no all-reduce per call); on exit the arrays registered with ``add`` are summed
    over the ranks IN PLACE with one all-reduce.  Sums commute with the np.diff the callers apply, and integer sums
    are exact, so the results are those of one all-reduce per call."""

    def __init__(self):
        self.pending = []

    def add(self, array):
        self.pending.append(array)
        return array

    def __enter__(self):
        _state["local"] += 1
        return self

    def __exit__(self, exc_type, exc, tb):
        _state["local"] -= 1
        if _state["local"] == 0 and _rank_world()[1] > 1:
            # a rank that failed inside the block must not leave its peers waiting in the all-reduce: agree on the
            # outcome first (one more tiny collective per statistic), then every rank raises
            failed = int(allreduce_sum(np.array([1 if exc_type is not None else 0], dtype=np.int64))[0])
            if failed and exc_type is None:
                raise RuntimeError("halotools_b200.distributed: %d peer rank(s) failed inside this statistic" % failed)
        if exc_type is None and self.pending and _state["local"] == 0:
            arrays, seen = [], set()
            for a in self.pending:
                if id(a) not in seen:
                    seen.add(id(a))
                    arrays.append(a)
            for dtype in sorted(set(a.dtype for a in arrays), key=str):      # the same order on every rank
                group = [a for a in arrays if a.dtype == dtype]
                flat = allreduce_sum(np.concatenate([a.ravel() for a in group]))
                pos = 0
                for a in group:
                    a[...] = flat[pos:pos + a.size].reshape(a.shape)
                    pos += a.size
        return False


def allreduce_sum(array):
    """Sum a small numpy array over the ranks (NCCL when the GPUs are there, gloo otherwise)."""
    rank, world = _rank_world()
    if world <= 1 or _state["local"] > 0:
        return array
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(_state["group"])
    t = torch.from_numpy(np.array(array, copy=True, order="C"))       # never reduce into the caller's array
    if backend == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_state["group"])
    return t.cpu().numpy()
