#!/bin/bash
# round 2: dense draw tuning pass (tests of the dense kernels, timing, ncu source-level profile of the blocked kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02b
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dense_blocked.py tests/test_gpu_regression_kernels.py tests/test_gpu_mcmc_regression.py -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
for p in 64 128 256; do
  timeout 300 python tools/perf_dense_draw.py 4096 $p >> $OUT/perf_dense.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_blocked -s 4 -c 1 -o $OUT/ncu_blocked64 python tools/perf_dense_draw.py 4096 64 > $OUT/ncu.log 2>&1
timeout 600 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -n 3 $OUT/pytest.log; cat $OUT/perf_dense.log; head -c 600 $OUT/bench_c2.json
