#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02o
mkdir -p $OUT
OMC_BENCH_PROFILE=1 timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02o/bench_default.json"))
print(d["value"], d["e2e"]["value"], d["e2e"]["phases_s"]); c3=d["c3"]; print(c3["value"], c3["e2e"]["value"], c3["e2e"]["phases_s"])
PY
grep -n "function calls" -A32 $OUT/bench_default.err | cut -c1-170 | tail -80
