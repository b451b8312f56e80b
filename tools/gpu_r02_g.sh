#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02g
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
for p in 64 56 48; do
  timeout 300 python tools/perf_dense_draw.py 4096 $p >> $OUT/perf_dense.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -n 4 $OUT/pytest.log; cat $OUT/perf_dense.log; head -c 300 $OUT/bench_c2.json
