#!/bin/bash
# Sweep library variants (variants/lib_*.so) and refinement overrides with a short resident bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for lib in variants/lib_*.so; do
  for m in ${SWEEP_M2:-default}; do
    if [ "$m" != "default" ]; then export HTB_M2=$m; else unset HTB_M2; fi
    echo "== $lib M2=$m M1=${HTB_M1:-default}"
    HTB_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln)
        print('ms_step %.1f  rr_ms %.1f  frac %.3f  evaluated %.3g value %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['pairs_evaluated_per_step'], d['value']), [ (c['ms_count'], c['tiles_redone']) for c in d['calls']])
    elif 'Error' in ln or 'error' in ln: print(ln.strip())
"
  done
done
