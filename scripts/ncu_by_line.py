"""Join an ncu SASS source page (per-instruction samples / executed counts) with nvdisasm line info of the
in-tree object file, and aggregate by source line.
usage: ncu_by_line.py <report.ncu-rep> <object.o> <kernel-symbol-substring> [top_n]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, obj, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur = None
infn = False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        infn = sym in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[1]
ia, ie, ism, ith = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), int(r[ie]), int(r[ism]), int(r[ith])))
    except Exception:
        pass
base = data[0][0]
agg = collections.defaultdict(lambda: [0, 0, 0])
for a, e, s, t in data:
    key = line_of.get(a - base, (("?", 0), ""))[0]
    g = agg[key]
    g[0] += e; g[1] += s; g[2] += t
te = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = {}
def text(f, n):
    if f not in src:
        for d in ("halotools_b200/csrc", "."):
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, f)
            if os.path.exists(p):
                src[f] = open(p).read().splitlines(); break
        else:
            src[f] = []
    L = src[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
print("# share of executed warp-instructions / of stall samples / avg active lanes, by source line (innermost inlined location)")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    f, n = key if key else ("?", 0)
    print("%5.2f%% inst %5.2f%% smp  lanes %4.1f  %s:%d  %s" % (100 * v[0] / te, 100 * v[1] / ts, v[2] / max(v[0], 1), f, n, text(f, n)))
