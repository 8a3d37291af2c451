"""Turn gpurun_out/prof_*.ncu-rep + launches.csv into the small text summaries kept under profiles/."""
import csv
import subprocess
import sys

rep, launches, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
lines = ["# ncu --set full --clock-control none, one launch of the dominant kernel (numbers under the profiler are NOT bench values)"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        lines.append("%-90s %-12s %s" % (k, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
sh = srows[1]
ia, ie, im = sh.index("Source"), sh.index("Instructions Executed"), sh.index("# Samples")
data = []
for r in srows[2:]:
    try:
        data.append((r[ia].strip(), int(r[ie]), int(r[im])))
    except Exception:
        pass
tot = sum(d[1] for d in data) or 1
mn = {}
for s, e, m in data:
    op = s.replace("@", " ").split()
    op = [t for t in op if not t.startswith(("P", "!P", "UP", "!UP"))]
    name = op[0].split(".")[0] if op else "?"
    mn[name] = mn.get(name, 0) + e
lines.append("")
lines.append("# executed warp-instructions by SASS mnemonic (share of %d)" % tot)
for k, v in sorted(mn.items(), key=lambda kv: -kv[1])[:24]:
    lines.append("%-12s %6.2f%%" % (k, 100.0 * v / tot))
lines.append("")
lines.append("# launch list (ncu --metrics gpu__time_duration.sum): kernel, launches, total ms, share")
agg = {}
try:
    lr = list(csv.reader(open(launches)))
    start = next(i for i, r in enumerate(lr) if r and r[0] == "ID")
    h = lr[start]
    ik, iv, imn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    for r in lr[start + 2:]:
        if len(r) > iv and r[imn] == "gpu__time_duration.sum":
            name = r[ik].split("(")[0]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += float(r[iv].replace(",", "")) / 1e6
    tt = sum(a[1] for a in agg.values()) or 1
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-60s %5d %10.3f ms %6.2f%%" % (k[:60], a[0], a[1], 100 * a[1] / tt))
except Exception as e:
    lines.append("launch list unavailable: %r" % (e,))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
