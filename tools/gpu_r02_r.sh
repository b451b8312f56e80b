#!/bin/bash
# round 2, run R: ncu --set full of the tridiagonal draw's kernels (fp32-assisted normals) + gmrf tests on the shipped lib
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02r
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gmrf.py tests/test_gpu_gmrf_module.py tests/test_gpu_mcmc_gmrf.py -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/pytest.log
python tools/perf_tridiag.py > $OUT/perf.txt 2>&1; tail -1 $OUT/perf.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tg_aggregate|tg_solve" -s 6 -c 2 -o $OUT/tridiag_f32 -f \
  python tools/perf_tridiag.py > $OUT/ncu.log 2>&1
ncu -i $OUT/tridiag_f32.ncu-rep --page raw --csv > $OUT/tridiag_f32.raw.csv 2>/dev/null
ncu -i $OUT/tridiag_f32.ncu-rep --page source --csv --kernel-name regex:tg_aggregate > $OUT/agg.src.csv 2>/dev/null
ncu -i $OUT/tridiag_f32.ncu-rep --page source --csv --kernel-name regex:tg_solve > $OUT/solve.src.csv 2>/dev/null
ls -la $OUT
