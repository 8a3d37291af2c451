cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== with sampler"; HTB_BENCH_VERBOSE=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep -o "step wall [0-9.]* ms\|\"ms_per_step\": [0-9.]*" | tr '\n' ' '
echo; echo "== without sampler"; HTB_BENCH_NO_SAMPLER=1 HTB_BENCH_VERBOSE=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | grep -o "step wall [0-9.]* ms\|\"ms_per_step\": [0-9.]*" | tr '\n' ' '
