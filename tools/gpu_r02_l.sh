#!/bin/bash
# round 2: full suite, default bench line, per-config lines, sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench default rc=$?" | tee -a $OUT/summary.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "bench reference rc=$?" | tee -a $OUT/summary.txt
for w in c1 c3 c4a c4b c5 c5full; do
  timeout 900 python bench.py --workload $w --steps 60 --warmup 5 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  echo "bench $w rc=$?" | tee -a $OUT/summary.txt
done
bash tools/gpu_sanitizer.sh > $OUT/sanitizer_stdout.log 2>&1
tail -n 3 $OUT/pytest.log; cat gpurun_out/sanitizer/summary.txt
