#!/bin/bash
# Round-1 measurement recipe (run under gpurun): plain bench of every workload first, then the ncu launch lists of the
# same commands, then one --set full capture of the top kernels.  Outputs under gpurun_out/; summaries are written into
# profiles/ afterwards with tools/ncu_summary.py.
set -x
mkdir -p gpurun_out
for w in c2 c1 c3 c4a c4b c5; do
  timeout 600 python bench.py --workload $w > gpurun_out/r01_bench_$w.json 2> gpurun_out/r01_bench_$w.err || echo "bench $w failed"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/r01_bench_c2.err
for w in c2 c3 c4a c4b c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tg_aggregate|tg_tilescan|tg_solve" -c 3 -s 9 -o gpurun_out/r01_tridiag -f \
  python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reg_pass -c 1 -s 3 -o gpurun_out/r01_reg_pass_full -f \
  python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rj_kernel" -c 3 -s 6 -o gpurun_out/r01_rj -f \
  python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mmala|random_walk" -c 1 -s 3 -o gpurun_out/r01_mmala -f \
  python bench.py --workload c4a --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_c4a.log 2>&1
# RandomWalkLoop (c4b) full capture as well
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"random_walk" -c 1 -s 3 -o gpurun_out/r01_rwl -f \
  python bench.py --workload c4b --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_c4b.log 2>&1
# the reports (with sources) exceed what gpurun brings back: export the raw pages here, keep only the CSVs
for r in gpurun_out/r01_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
  rm -f $r
done
rm -f gpurun_out/ncu_*.log
ls -la gpurun_out | tail -40
du -sh gpurun_out
