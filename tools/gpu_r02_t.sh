#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02t
mkdir -p $OUT
OMC_BENCH_PROFILE=1 timeout 600 python bench.py --workload c3 --no-cpu --no-extras > $OUT/bench_c3.json 2> $OUT/bench_c3.err
grep -n "function calls" -A45 $OUT/bench_c3.err | cut -c1-180 | head -70
