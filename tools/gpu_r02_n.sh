#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02n
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_rj.py -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
CUDA_LAUNCH_BLOCKING=1 timeout 600 python tools/profile_prepare.py c3 > $OUT/prof_c3_blocking.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rj_kernel -s 6 -c 3 -o $OUT/ncu_rj python bench.py --workload c5 --steps 6 --warmup 2 --no-cpu --no-extras --no-e2e > $OUT/ncu.log 2>&1
tail -n 4 $OUT/pytest.log
grep -n "== c3" -A48 $OUT/prof_c3_blocking.log | tail -52 | cut -c1-160
