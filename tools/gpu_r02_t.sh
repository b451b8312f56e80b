#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02t
mkdir -p $OUT
W=${1:-c3}
OMC_BENCH_PROFILE=1 timeout 600 python bench.py --workload $W --no-cpu --no-extras > $OUT/bench_$W.json 2> $OUT/bench_$W.err
grep -n "function calls" -A45 $OUT/bench_$W.err | cut -c1-180 | head -70
