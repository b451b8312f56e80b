#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02y
mkdir -p $OUT
for prep in stat empty warm; do
  OMC_BENCH_E2E_PREP=$prep timeout 900 python bench.py --no-cpu > $OUT/bench_$prep.json 2> $OUT/bench_$prep.err
  grep "allocator state" $OUT/bench_$prep.err
  python - <<PY
import json
d=[json.loads(l) for l in open("$OUT/bench_$prep.json") if l.startswith("{")][-1]
e=d["e2e"]; print("$prep", e["value"], e["seconds"], e["phases_s"], [(b["queue_next_s"], b["plan_s"]) for b in e["blocks"]][:5])
PY
done
