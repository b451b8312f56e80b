#!/bin/bash
# round 2, run S: full GPU suite, the bench lines of every workload, the reference arm, ncu launch lists and --set full
# captures of the C2 sweep and the C3 draw (each only after its own command ran clean without ncu).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02s
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench reference rc=$?"
for w in c1 c3 c4a c4b c5 c5full; do
  timeout 600 python bench.py --workload $w --no-extras > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"
done
for w in c2 c3 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $OUT/ncu_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nn_warp_draw|fused_small" -s 8 -c 3 -o $OUT/c2_full -f \
  python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $OUT/ncu_full_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tg_aggregate|tg_solve|tg_tilescan|tg_partials" -s 8 -c 4 -o $OUT/c3_full -f \
  python tools/perf_tridiag.py > $OUT/ncu_full_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rj_kernel" -s 6 -c 3 -o $OUT/c5_full -f \
  python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $OUT/ncu_full_c5.log 2>&1
for r in $OUT/*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
done
ls -la $OUT | tail -40
