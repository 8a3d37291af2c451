#!/usr/bin/env python
"""Benchmark of the pair-counting hot path (BASELINE.json: "pair evals/sec (GPairs/s)").

Workload (N = 1 and every N): BASELINE.json configs[1] — tpcf Landy-Szalay on a zheng07 HOD mock
populated on FakeSim-style synthetic halos (~5e5 galaxies) plus 5e6 uniform randoms, Lbox = 250,
15 log rbins 0.1-20: one STEP = the three pair counts the estimator needs (DD, DR, RR through
npairs_3d) + the estimator.  Pairs are counted in units of W_ref = the (i, j) pairs the REFERENCE
mesh loop visits for the same call, so the GPU arm and the reference arm quote the
same unit of work.

  value     whole-job GPairs/s with the samples already resident in HBM: hb.tpcf(device tensors) - mesh sorts,
            DD / RR / DR count kernels on three streams, ONE all-reduce of the count tables on the device,
            Landy-Szalay by the estimator kernel, one host synchronisation (the D2H of xi) per step
  e2e       the same through the public host API halotools_b200.tpcf(numpy in -> numpy out), pinned
            host arrays, H2D of every sample inside the timed region
  roofline  FP64 non-FMA issue roofline of the dominant kernel (k_count<Fast3>): 8 f64 ops per
            evaluated pair (3 sub, 3 mul, 2 add) x pairs evaluated / kernel time, against the FP64
            DADD/DMUL issue rate measured live on the same GPU (MEASURED_PEAKS.json has no FP64 entry).
            The step's three launches (DD, DR, RR) run co-resident on three streams, so the time is the
            window in which they ran inside the timed steps (first warp in of any -> last warp out of
            any, from stamps the kernels write themselves) and the pairs are those of all three; the RR
            launch run alone is reported beside it (roofline.rr_launch)
  e2e_pageable   the same from ordinary (pageable) numpy arrays - what a drop-in caller passes
  cpu_baseline / --impl reference   the reference's own compiled Cython engine (oracle/_ref) — or the
            C oracle port when it is absent — on all host cores, on a bounded sample (a range of
            mesh1 cells of the RR count); engine time only, the mesh build is reported beside it.
  parity    (N = 1) the step's counts against the reference's engine AT FULL SIZE: DD in full, DR and RR
            on the mesh1 cell range the CPU sample covers; the bench fails if any count differs
  c5        sub-record: BASELINE configs[4] (mean_delta_sigma, 1e6 x 1e8), the north-star scaling config,
            measured in the same run at the same N (same keys as the main line)

Multi-GPU (torchrun, one rank per GPU): every rank holds both samples, counts ITS work-balanced contiguous
range of reference mesh1 cells, one NCCL all-reduce of the count tables per step (on the device, on the engine's
stream); the end-to-end arm uploads 1/N of every sample per rank and all-gathers over NVLink; total work is
fixed -> "strong" scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LBOX = 250.0
N_RANDOMS = 5_000_000
OPS_PER_PAIR = 8.0


def make_inputs(n_randoms=N_RANDOMS):
    from halotools_b200 import synthetic
    gal = synthetic.fakesim_zheng07_mock(560, LBOX, seed=43)
    ran = synthetic.uniform_points(44, n_randoms, LBOX)
    return gal, ran, synthetic.config_rbins()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi takes a while to attach to the driver; do not let that overlap the timed region
            t0 = time.time()
            while not self.lines and time.time() - t0 < 10.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def landy_szalay(DD, DR, RR, N, NR):
    from halotools_b200.two_point_clustering.tpcf_estimators import _TP_estimator
    return _TP_estimator(np.diff(DD), np.diff(DR), np.diff(RR), N, N, NR, NR, "Landy-Szalay")


# ------------------------------------------------------------------ reference / CPU arm
class CpuArm(object):
    """The reference's compiled npairs_3d engine (oracle/_ref; the C oracle port when it did not travel) on all host
    cores, on a bounded sample of the step: the RR count restricted to a contiguous range of mesh1 cells.  The double
    mesh and the worker pool are built ONCE, outside the timed region (a per-call set-up spread over 2 % of
    the cells understates the engine); their cost is reported separately (``seconds_mesh``)."""

    def __init__(self, ran, rbins, cores, target_cells_per_core=2):
        from oracle import oracle, ref_engines
        self.ran, self.rbins, self.cores = ran, rbins, cores
        self.kind = "reference" if ref_engines.available() else "port"
        if self.kind == "reference":
            self.prep = ref_engines.PreparedCount(ran, ran, rbins, LBOX, cores)
            self.dm = self.prep.dm
            self.seconds_mesh = self.prep.seconds_mesh
        else:
            t0 = time.perf_counter()
            self.dm = oracle.build_double_mesh_3d(ran, ran, [float(rbins.max())] * 3, LBOX, None, None)[0]
            self.seconds_mesh = time.perf_counter() - t0
            self.prep = None
        ncells = self.dm.mesh1.ncells
        self.cells = (0, int(min(ncells, max(cores * target_cells_per_core, 8))))
        self.pairs = float(self.dm.visited_pairs(*self.cells))
        self.sample = ("RR count of the %d randoms restricted to mesh1 cells [%d, %d) of %d (%.3g visited pairs), num_threads=%d; "
                       "engine time only - the double mesh (%.2f s, serial numpy argsort as in the reference) and the worker "
                       "pool are built once outside the timed region"
                       % (len(ran), self.cells[0], self.cells[1], ncells, self.pairs, cores, self.seconds_mesh))

    def run(self):
        """-> (GPairs/s, seconds, counts) of one pass over the sample."""
        from oracle import oracle
        t0 = time.perf_counter()
        if self.prep is not None:
            counts = self.prep.run(self.cells)
        else:
            counts = oracle.npairs_3d(self.ran, self.ran, self.rbins, period=LBOX, num_threads=self.cores, cell1_range=self.cells)
        dt = time.perf_counter() - t0
        return self.pairs / dt / 1e9, dt, np.asarray(counts)

    def close(self):
        if self.prep is not None:
            self.prep.close()


WORKLOAD = ("configs[1]: tpcf Landy-Szalay (DD+DR+RR via npairs_3d), zheng07-on-FakeSim mock + 5e6 randoms, Lbox 250, "
            "15 log rbins 0.1-20")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    gal, ran, rbins = make_inputs(args.randoms)
    cores = os.cpu_count() or 1
    arm = CpuArm(ran, rbins, cores)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, dt, _ = arm.run()
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    arm.close()
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "pair evals/sec", "value": value, "unit": "GPairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD + " (each step = a bounded sample of the RR count, rate in the same W_ref pair unit)",
                       "same_config": False,
                       "method": "rate of the reference engine on a cell range of the dominant (RR) count; the GPU arm's "
                                 "value is W_ref of the whole step / its time"},
            "cpu_baseline": {"value": value, "unit": "GPairs/s", "cores": cores, "kind": arm.kind, "sample": arm.sample,
                             "seconds_mesh": arm.seconds_mesh},
            "e2e": {"value": value, "unit": "GPairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ GPU arm: shared plumbing
class Env(object):
    def __init__(self):
        import ctypes
        import torch
        import torch.distributed as dist
        from halotools_b200 import _lib, distributed
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        _lib.require_gpu()
        torch.cuda.set_device(self.local)
        _lib.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            distributed.enable()
        self.stream = torch.cuda.Stream()
        _lib.check(_lib.load().htb_set_stream(ctypes.c_void_p(self.stream.cuda_stream)))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed calls, then EXACTLY `steps` calls between two CUDA events on the engine's stream, bracketed by a
        barrier + synchronize on both sides; max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            res = None
            for _ in range(steps):
                res = fn()
            e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, res

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


# ------------------------------------------------------------------ configs[1]: the tpcf step
def measure_tpcf(env, args, sampler):
    import halotools_b200 as hb
    from halotools_b200 import _lib, distributed
    torch, dist, world, rank = env.torch, env.dist, env.world, env.rank

    gal, ran, rbins = make_inputs(args.randoms)
    N, NR = len(gal), len(ran)
    # pinned host copies for the end-to-end arm; the ordinary (pageable) numpy arrays for e2e_pageable
    gal_h = torch.from_numpy(gal).pin_memory()
    ran_h = torch.from_numpy(ran).pin_memory()
    gal_np, ran_np = gal_h.numpy(), ran_h.numpy()
    gal_d, ran_d = gal_h.cuda(non_blocking=False), ran_h.cuda(non_blocking=False)
    torch.cuda.synchronize()
    stats_acc = {}

    def stats_pass():
        """UNTIMED: the step's three counts as synchronous engine calls, for the work counters (W_ref, evaluated pairs),
        the CUDA-event time of every count kernel (roofline) and the launch count.  The timed step below runs the same
        kernels on the same inputs through hb.tpcf."""
        acc = {"pairs_reference": 0.0, "pairs_evaluated": 0.0, "ms_count": 0.0, "launches": 0, "ms_mesh": 0.0,
               "count_evaluated": [], "count_ms": [], "calls": []}
        out = []
        with distributed.local_counts() as part:
            for a, b in ((gal_d, gal_d), (gal_d, ran_d), (ran_d, ran_d)):
                out.append(part.add(hb.npairs_3d(a, b, rbins, period=LBOX)))
                st = _lib.last_stats
                acc["pairs_reference"] += st["pairs_reference"]
                acc["pairs_evaluated"] += st["pairs_evaluated"]
                acc["ms_count"] += st["ms_count"]
                acc["ms_mesh"] += st["ms_mesh"]
                acc["launches"] += st["kernel_launches"] - 1          # (the W_ref sum kernel only runs with stats)
                acc["count_evaluated"].append(st["pairs_evaluated"])
                acc["count_ms"].append(st["ms_count"])
                acc["calls"].append({k: st[k] for k in ("ms_h2d", "ms_mesh", "ms_count", "ms_total", "tiles",
                                                        "tiles_redone", "refine1", "refine2")})
        acc["launches"] += 1                                          # + the estimator kernel of the timed step
        stats_acc.update(acc)
        return landy_szalay(out[0], out[1], out[2], N, NR)

    _lib.async_count_times()          # (reset the event ring)

    def step_resident():
        # device tensors in: DD, DR, RR enqueued back to back on one stream, ONE all-reduce of the count tables on the
        # device, Landy-Szalay by the estimator kernel, one host synchronisation (the D2H of xi)
        return hb.tpcf(gal_d, rbins, randoms=ran_d, period=LBOX, estimator="Landy-Szalay")

    def step_e2e():
        return hb.tpcf(gal_np, rbins, randoms=ran_np, period=LBOX, estimator="Landy-Szalay")

    def step_e2e_pageable():
        return hb.tpcf(gal, rbins, randoms=ran, period=LBOX, estimator="Landy-Szalay")

    stats_pass()                       # (first calls: module load, pool growth, function attributes)
    xi_host = stats_pass()
    if sampler is not None:
        sampler.start()
    ms_step, xi_res = env.timed(step_resident, args.steps, args.warmup)
    acc = dict(stats_acc)
    # CUDA-event durations of the count kernels INSIDE the timed region (three per step; the RR launch is the longest)
    # ... and the kernels' own device-side stamps (first warp in -> last warp out): the three counts run on three streams,
    # so an event bracket around the RR launch also holds the time its blocks waited for DD / DR blocks to retire
    ks = _lib.async_kernel_spans()
    stamps = _lib.async_kernel_stamps()
    kt = _lib.async_count_times()
    kt = kt[len(kt) % 3:]
    ks = ks[len(ks) % 3:]
    stamps = stamps[len(stamps) % 3:]
    rr_events = [max(kt[i:i + 3]) for i in range(0, len(kt), 3)]
    rr_spans = [max(ks[i:i + 3]) for i in range(0, len(ks), 3)]
    # the three launches of a step are co-resident (three streams): the window in which they ran is the union of their spans
    windows = []
    for i in range(0, len(stamps), 3):
        trio = [t for t in stamps[i:i + 3] if t[0] and t[1]]
        if len(trio) == 3:
            windows.append((max(t[1] for t in trio) - min(t[0] for t in trio)) * 1e-6)
    acc["rr_ms_events"] = float(np.mean(rr_events)) if rr_events else None
    acc["rr_ms_span"] = float(np.mean(rr_spans)) if rr_spans and min(rr_spans) > 0 else None
    acc["count_window_ms"] = float(np.mean(windows)) if windows else None
    for k in ("rr_ms_events", "rr_ms_span", "count_window_ms"):
        stats_acc[k] = acc[k]
    # the estimator kernel evaluates the reference's numpy expressions operation by operation
    assert np.allclose(xi_res, xi_host, rtol=1e-14, atol=0), "device estimator and host Landy-Szalay disagree"
    if world > 1:
        # per-rank stats describe this rank's shard; totals over ranks
        t = torch.tensor([acc["pairs_reference"], acc["pairs_evaluated"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        acc["pairs_reference"], acc["pairs_evaluated"] = float(t[0]), float(t[1])
    ms_e2e, xi_e2e = env.timed(step_e2e, max(1, args.steps // 2), 1)
    ms_pg, xi_pg = env.timed(step_e2e_pageable, max(1, args.steps // 2), 1)
    clocks = sampler.stop() if sampler is not None else None
    assert np.allclose(xi_res, xi_e2e, rtol=1e-12, atol=0), "resident and end-to-end paths disagree"
    assert np.allclose(xi_res, xi_pg, rtol=1e-12, atol=0), "resident and pageable end-to-end paths disagree"

    W = acc["pairs_reference"]
    out = {"ms_step": ms_step, "W": W, "N": N, "NR": NR, "acc": acc, "clocks": clocks, "stats": stats_acc,
           "ms_e2e": ms_e2e, "ms_e2e_pageable": ms_pg, "rbins": rbins}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # ---- reference arm beside the GPU arm, and REFERENCE PARITY AT FULL SIZE: the reference's own
        # engine on the same inputs - DD in full, DR and RR on the cell range the CPU sample times
        from oracle import oracle, ref_engines
        cores = os.cpu_count() or 1
        arm = CpuArm(ran, rbins, cores)
        v, dt, rr_ref = arm.run()
        arm.close()
        out["cpu"] = {"value": v, "unit": "GPairs/s", "cores": cores, "kind": arm.kind, "sample": arm.sample,
                      "seconds": dt, "seconds_mesh": arm.seconds_mesh}
        REF = ref_engines if ref_engines.available() else oracle
        cells = arm.cells
        with distributed.cell_range(*cells):
            rr_gpu = hb.npairs_3d(ran_d, ran_d, rbins, period=LBOX)
            dr_gpu = hb.npairs_3d(gal_d, ran_d, rbins, period=LBOX)
        dr_ref = REF.npairs_3d(gal, ran, rbins, period=LBOX, num_threads=cores, cell1_range=cells)
        dd_ref = REF.npairs_3d(gal, gal, rbins, period=LBOX, num_threads=cores)
        dd_gpu = hb.npairs_3d(gal_d, gal_d, rbins, period=LBOX)
        ok = {"DD_full": bool(np.array_equal(dd_gpu, dd_ref)), "DR_cells": bool(np.array_equal(dr_gpu, dr_ref)),
              "RR_cells": bool(np.array_equal(rr_gpu, rr_ref))}
        out["parity"] = {"against": arm.kind, "mesh1_cells": list(cells), "bit_exact": ok,
                         "DD_top": int(dd_gpu[-1]), "DR_top_cells": int(dr_gpu[-1]), "RR_top_cells": int(rr_gpu[-1])}
        assert all(ok.values()), "full-size counts differ from the reference: %r" % (out["parity"],)
    return out


def tpcf_line(env, args, m):
    from halotools_b200 import _lib
    world = env.world
    acc, stats_acc, W, N, NR = m["acc"], m["stats"], m["W"], m["N"], m["NR"]
    # FP64 issue-rate roofline of the dominant kernel (the RR launch of k_count<Fast3>) on this rank
    rate, clk = _lib.measure_fp64_rate()
    i_rr = int(np.argmax(stats_acc["count_evaluated"]))
    # The three count launches of a step (DD, DR, RR: all k_count<Fast3>) run on three streams and share the SMs, so the
    # RR launch alone has no duration of its own inside a step: the roofline is taken over the three of them - their
    # algorithmic FP64 operations over the window in which they ran (first warp in of any -> last warp out of any, from
    # the kernels' device-side stamps, averaged over the timed steps).  The RR launch run ALONE (the synchronous
    # statistics pass, same inputs) is reported beside it.
    pairs_all = float(sum(stats_acc["count_evaluated"]))
    window_ms = stats_acc.get("count_window_ms")
    alone_ms = stats_acc["count_ms"][i_rr]
    frac_alone = stats_acc["count_evaluated"][i_rr] * OPS_PER_PAIR / (alone_ms * 1e-3) / rate
    if window_ms:
        kernel_ms, pairs_k = window_ms, pairs_all
        what = "k_count<Fast3>: the DD + DR + RR launches of a step (three streams, co-resident)"
    else:
        kernel_ms, pairs_k = alone_ms, stats_acc["count_evaluated"][i_rr]
        what = "k_count<Fast3> (RR launch, statistics pass)"
    ach = pairs_k * OPS_PER_PAIR / (kernel_ms * 1e-3) / 1e12
    peak = rate / 1e12
    traffic = None
    for name in ("r02_fast3_traffic.json", "r01_fast3_traffic.json"):
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel, from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                tj = json.load(fh)
            traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
            break
        except Exception:
            traffic = None
    roofline = {"bound": "fp64_issue", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": traffic, "traffic_unit": "bytes per RR launch (ncu --set full capture under profiles/)",
                "kernel": what,
                "peak_source": "htb_measure_fp64_rate: DADD/DMUL non-FMA issue rate measured live on this GPU "
                               "(MEASURED_PEAKS.json has no FP64 entry)",
                "pairs_evaluated_per_launch": pairs_k,
                "kernel_ms": kernel_ms,
                "kernel_ms_source": "device-side globaltimer stamps written by the kernels (first warp in -> last warp out), union "
                                    "over the step's three launches, averaged over the timed steps",
                "rr_launch": {"pairs_evaluated": stats_acc["count_evaluated"][i_rr], "ms_alone": alone_ms, "frac_alone": frac_alone,
                              "ms_span_in_step": stats_acc.get("rr_ms_span"), "ms_event_bracket_in_step": stats_acc.get("rr_ms_events"),
                              "note": "inside a step the RR launch shares the SMs with the DR and DD launches for its whole "
                                      "length (surplus blocks retire at once, HTB_FLAG_EARLY_EXIT), so its span there is longer "
                                      "than when it runs alone"}}
    h2d = int((N + NR) * 24)          # inside hb.tpcf the upload cache sends every sample across PCIe once per step
    d2h = int(3 * len(m["rbins"]) * 8)
    line = {"metric": "pair evals/sec", "value": W / (m["ms_step"] * 1e-3) / 1e9, "unit": "GPairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD + " (%d galaxies, %d randoms)" % (N, NR),
                       "pairs_unit": "W_ref = pairs visited by the reference mesh loop for the same calls",
                       "pairs_reference_per_step": W, "pairs_evaluated_per_step": acc["pairs_evaluated"],
                       "l2": "inputs (%.0f MB of sorted coordinates) exceed nothing that matters: the kernel is "
                             "FP64-issue bound; every step re-sorts both samples and re-streams them from HBM"
                             % ((N + NR) * 24 / 1e6),
                       "parallelism": "work-balanced mesh1 cell ranges over %d rank(s), one NCCL all-reduce of the count tables per "
                                      "step (device buffer, engine stream); e2e: 1/N upload per rank + all-gather" % world,
                       "streams": "the three counts of a step run on three CUDA streams (tail of one persistent kernel filled by "
                                  "the next); the roofline is taken over the three launches together"},
            "e2e": {"value": W / (m["ms_e2e"] * 1e-3) / 1e9, "unit": "GPairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": m["ms_e2e"], "host_memory": "pinned"},
            "e2e_pageable": {"value": W / (m["ms_e2e_pageable"] * 1e-3) / 1e9, "unit": "GPairs/s", "h2d_bytes_per_step": h2d,
                             "d2h_bytes_per_step": d2h, "ms_per_step": m["ms_e2e_pageable"],
                             "host_memory": "pageable (ordinary numpy arrays, what a drop-in caller passes)"},
            "gpu_launches": int(acc["launches"] * args.steps),
            "roofline": roofline, "cpu_baseline": m.get("cpu"), "parity": m.get("parity"), "clocks": m["clocks"],
            "breakdown_ms": {"mesh_sort": acc["ms_mesh"], "count_kernels": acc["ms_count"]},
            "calls": acc.get("calls")}
    return line


def run_gpu(args):
    env = Env()
    sampler = ClockSampler(env.local) if (env.rank == 0 and not os.environ.get("HTB_BENCH_NO_SAMPLER")) else None
    m = measure_tpcf(env, args, sampler)
    c5 = None
    if not args.no_c5:
        # the north-star scaling config rides on the driver's line at every N
        c5 = measure_c5(env, args, None, steps=max(2, min(args.steps, 3)), warmup=1)
    if env.rank == 0:
        line = tpcf_line(env, args, m)
        line["c5"] = c5
        print(json.dumps(line))
    env.close()
    return 0


# ------------------------------------------------------------------ configs[4]: the delta-sigma scaling config
def measure_c5(env, args, sampler, steps=None, warmup=None):
    """BASELINE.json configs[4]: mean_delta_sigma, 1e6 galaxies x 1e8 particles, Lbox 1000, 15 log rp bins 0.1-30,
    one particle mass.  Same JSON contract; W_ref = pairs the reference's 2-d mesh loop visits."""
    import halotools_b200 as hb
    from halotools_b200 import _lib, synthetic
    torch, dist, world, rank = env.torch, env.dist, env.world, env.rank
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    ngal, nptcl, L = args.c5_galaxies, args.c5_particles, 1000.0
    gal = synthetic.uniform_points(43, ngal, L)
    ptcl = synthetic.uniform_points(44, nptcl, L)
    rp = np.logspace(-1, np.log10(30.0), 15)
    gal_h, ptcl_h = torch.from_numpy(gal).pin_memory(), torch.from_numpy(ptcl).pin_memory()
    del gal, ptcl
    gal_np, ptcl_np = gal_h.numpy(), ptcl_h.numpy()
    gal_d, ptcl_d = gal_h.cuda(), ptcl_h.cuda()
    torch.cuda.synchronize()
    acc = {}

    def stats_pass():
        # UNTIMED: the same call with the ranks' results left un-reduced, i.e. as a synchronous engine call that fills the
        # work counters and the CUDA-event time of the count kernel (multi-GPU steps leave their sums on the device and
        # return no per-call statistics)
        from halotools_b200 import distributed
        with distributed.local_counts():
            hb.mean_delta_sigma(gal_d, ptcl_d, 1.0, rp, period=L)
            acc.update(_lib.last_stats)

    def step_resident():
        return hb.mean_delta_sigma(gal_d, ptcl_d, 1.0, rp, period=L)

    def step_e2e():
        return hb.mean_delta_sigma(gal_np, ptcl_np, 1.0, rp, period=L)

    stats_pass()
    stats_pass()
    if sampler is not None:
        sampler.start()
    ms_step, res = env.timed(step_resident, steps, warmup)
    st = dict(acc)
    tot = torch.tensor([st["pairs_reference"], st["pairs_evaluated"], st["ms_count"]], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        st["ms_count"] = float(mx[2])
    W, Wgpu = float(tot[0]), float(tot[1])
    ms_e2e, res2 = env.timed(step_e2e, max(1, steps // 2), 1)
    clocks = sampler.stop() if sampler is not None else None
    scale = float(np.max(np.abs(res)))
    assert np.allclose(res, res2, rtol=1e-9, atol=1e-12 * scale), "resident and end-to-end paths disagree"
    del gal_d, ptcl_d, gal_h, ptcl_h
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    rate, _ = _lib.measure_fp64_rate()
    ach = Wgpu / world * 5.0 / (st["ms_count"] * 1e-3) / 1e12
    return {"metric": "pair evals/sec", "value": W / (ms_step * 1e-3) / 1e9, "unit": "GPairs/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[4]: mean_delta_sigma, %d galaxies x %d particles, Lbox 1000, 15 log rp bins "
                                   "0.1-30, one particle mass" % (ngal, nptcl),
                       "pairs_unit": "W_ref = pairs visited by the reference mesh loop for the same call",
                       "pairs_reference_per_step": W, "pairs_evaluated_per_step": Wgpu,
                       "l2": "inputs (%.1f GB of coordinates) are larger than L2" % ((ngal + nptcl) * 24 / 1e9),
                       "parallelism": "work-balanced mesh1 cell ranges over %d rank(s), one all-reduce of the column sums" % world},
            "e2e": {"value": W / (ms_e2e * 1e-3) / 1e9, "unit": "GPairs/s", "h2d_bytes_per_step": int((ngal + nptcl) * 16),    # the x and y columns (the engine never reads z)
                    "d2h_bytes_per_step": int(14 * 8), "ms_per_step": ms_e2e},
            "gpu_launches": int(st["kernel_launches"] * steps),
            "roofline": {"bound": "fp64_issue", "achieved": ach, "peak": rate / 1e12, "unit": "TFLOP/s", "frac": ach / (rate / 1e12),
                         "traffic": None, "kernel": "k_count<DSigmaR> (5 f64 ops per evaluated pair; slowest rank)",
                         "kernel_ms": st["ms_count"]},
            "cpu_baseline": None, "clocks": clocks,
            "breakdown_ms": {"mesh_sort": st["ms_mesh"], "count_kernel": st["ms_count"], "path": st["path"]},
            "delta_sigma": [float(v) for v in res]}


def run_gpu_c5(args):
    env = Env()
    sampler = ClockSampler(env.local) if (env.rank == 0 and not os.environ.get("HTB_BENCH_NO_SAMPLER")) else None
    line = measure_c5(env, args, sampler)
    if env.rank == 0:
        print(json.dumps(line))
    env.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--randoms", type=int, default=N_RANDOMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the configs[4] sub-record of the default line")
    ap.add_argument("--workload", default="tpcf", choices=["tpcf", "c5"],
                    help="tpcf = BASELINE configs[1] (the driver's line); c5 = configs[4], the delta-sigma scaling config")
    ap.add_argument("--c5-galaxies", type=int, default=1_000_000)
    ap.add_argument("--c5-particles", type=int, default=100_000_000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        return run_gpu_c5(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
