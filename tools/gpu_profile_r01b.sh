#!/bin/bash
# Round-1 (second half) measurement recipe, run under gpurun after the C2 sweep changed (residual-only pass, column
# Cholesky): plain bench of every workload first, then the ncu launch list of the C2 command, then one --set full
# capture of its kernels.  Outputs under gpurun_out/; summaries go to profiles/ with tools/ncu_summary.py.
set -x
mkdir -p gpurun_out
for w in c2 c1 c3 c4a c4b c5 c5full; do
  timeout 600 python bench.py --workload $w > gpurun_out/r01_bench_$w.json 2> gpurun_out/r01_bench_$w.err || echo "bench $w failed"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/r01_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_c2.csv \
  python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reg_pass_kernel|nn_dense_draw" -c 4 -s 2 -o gpurun_out/r01b_c2_full -f \
  python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_c4a.csv \
  python bench.py --workload c4a --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c4a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_c5.csv \
  python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mmala" -c 1 -s 3 -o gpurun_out/r01b_mmala_diag -f \
  python bench.py --workload c4a --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_c4a.log 2>&1
for r in gpurun_out/r01b_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
  rm -f $r
done
rm -f gpurun_out/ncu_*.log
ls -la gpurun_out | tail -20
