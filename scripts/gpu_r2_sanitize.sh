#!/bin/bash
# compute-sanitizer evidence: memcheck over every golden case (all kernel families, the
# asynchronous statistics, the device input step) and over the kernel-variant flags of the integer engines; racecheck
# (shared-memory hazards: TMA staging ring, per-lane queues, BinQ lists) over one case per kernel family.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # tool, log name, pytest -k expression
  timeout ${4:-1500} compute-sanitizer --tool $1 --print-limit 3 --launch-timeout 600 --error-exitcode 77 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$3" > gpurun_out/$2.log 2>&1
  echo "$1 [$3] rc=$?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/$2.log | tail -4
}
run memcheck sanitize_memcheck_golden "test_matches_reference_golden and not n3d_c1_full and not tpcf_jk and not wp_jk and not rp_pi_jk"
run memcheck sanitize_memcheck_variants "test_kernel_variants_agree and (n3d_periodic or n3d_clustered or xyz_periodic or smu_periodic or perobj_periodic or proj_periodic)"
run memcheck sanitize_memcheck_device "device_estimator or asynchronous or input_step or device_statistics"
run racecheck sanitize_racecheck "test_matches_reference_golden and (n3d_periodic or n3d_clustered or xyz_wp_like or mxyz_id01 or marked_id01 or marked_id03 or ds_periodic_mean or ds_masses or smu_periodic or rp_pi_auto or perobj_periodic or jk3d_periodic or wxy_periodic or tpcf_randoms_Landy)" 2400
