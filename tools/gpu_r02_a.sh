#!/bin/bash
# round 2, first GPU pass: new dense kernels (blocked Cholesky, wide regression panels, re-centred rss)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02a
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dense_blocked.py tests/test_gpu_regression_kernels.py -x -q > $OUT/pytest_dense.log 2>&1
echo "pytest dense rc=$?" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_mcmc_regression.py tests/test_gpu_mh.py -x -q > $OUT/pytest_mcmc.log 2>&1
echo "pytest mcmc rc=$?" | tee -a $OUT/summary.txt
for p in 64 128 256; do
  timeout 300 python tools/perf_dense_draw.py 4096 $p >> $OUT/perf_dense.log 2>&1
done
OMC_DENSE_DRAW_IMPL=columns timeout 300 python tools/perf_dense_draw.py 4096 64 >> $OUT/perf_dense.log 2>&1
timeout 600 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_dense.log $OUT/pytest_mcmc.log; cat $OUT/perf_dense.log; cat $OUT/bench_c2.json | head -c 3000
