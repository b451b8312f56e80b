#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the small-shape parity runs of every kernel family; logs -> profiles/
# (SURVEY §5 "race detection": the reference is single-threaded; the kernels here share memory between warps and, in
# tridiag.cu, pass look-back records between CTAs through fence-free 16-byte flag+payload words).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/sanitizer
mkdir -p $OUT
SEL_TRIDIAG='tests/test_gpu_gmrf.py::test_mcmc_replays_reference_gmrf_chain tests/test_gpu_gmrf.py::test_tridiag_not_positive_definite_sets_status'
SEL_DENSE='tests/test_gpu_regression_kernels.py::test_nn_dense_draw_without_probes_matches_oracle tests/test_gpu_dense_blocked.py::test_dense_factor_flags_non_pd tests/test_gpu_fused_small.py'
SEL_MH='tests/test_gpu_mh.py::test_mmala_replays_reference_chain_analytic tests/test_gpu_mh.py::test_random_walk_loop_replays_reference_chain tests/test_gpu_mh.py::test_mmala_poisson_gamma_chains'
SEL_RJ='tests/test_gpu_rj.py::test_rj_kernel_replays_reference_steps tests/test_gpu_rj.py::test_rj_companion_samplers_replay_reference_calls'
for fam in TRIDIAG DENSE MH RJ; do
  sel=SEL_$fam
  for tool in memcheck racecheck; do
    log=$OUT/${tool}_$(echo $fam | tr A-Z a-z).log
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --log-file $log.raw \
      python -m pytest ${!sel} -x -q -p no:cacheprovider > $log.pytest 2>&1
    rc=$?
    { echo "# compute-sanitizer --tool $tool : ${!sel}"; echo "# exit code $rc (77 = sanitizer errors)"; tail -n 3 $log.pytest; \
      grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard" $log.raw | sort | uniq -c | head -20; } > $log
    rm -f $log.raw $log.pytest
    echo "$tool $fam rc=$rc" | tee -a $OUT/summary.txt
  done
done
cat $OUT/*.log | head -80
