#!/bin/bash
# Round-1 measurement recipe (run under gpurun): plain bench first, then the ncu launch lists of the same commands,
# then one --set full capture of the top kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || exit 1
tail -c 600 gpurun_out/bench_c2.json
for w in c2 c3 c4a c4b; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tridiag_forward -c 1 -s 3 -o gpurun_out/r01_tridiag_forward -f \
  python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tridiag_backward -c 1 -s 3 -o gpurun_out/r01_tridiag_backward -f \
  python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_bwd.log 2>&1
ls -la gpurun_out | tail -20
