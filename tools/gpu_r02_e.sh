#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02f
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_regression_kernels.py tests/test_gpu_dense_blocked.py tests/test_gpu_mcmc_regression.py tests/test_gpu_mh.py tests/test_gpu_stream_store.py tests/test_gpu_rj.py -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
for impl in warp columns blocked; do
  for p in 64 32 40; do
    OMC_DENSE_DRAW_IMPL=$impl timeout 300 python tools/perf_dense_draw.py 4096 $p >> $OUT/perf_dense.log 2>&1
  done
done
OMC_DENSE_DRAW_IMPL=warp timeout 300 python tools/perf_dense_draw.py 4096 3 >> $OUT/perf_dense.log 2>&1
OMC_DENSE_DRAW_IMPL=columns timeout 300 python tools/perf_dense_draw.py 4096 3 >> $OUT/perf_dense.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_warp -s 4 -c 1 -o $OUT/ncu_warp64 python tools/perf_dense_draw.py 4096 64 > $OUT/ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -n 4 $OUT/pytest.log; cat $OUT/perf_dense.log; head -c 400 $OUT/bench_c2.json
