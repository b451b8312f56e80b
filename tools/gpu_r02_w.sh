#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02w
mkdir -p $OUT
for r in 1 2 3; do
  timeout 600 python bench.py --workload c2 --no-cpu --no-extras > $OUT/bench_c2_r$r.json 2> $OUT/bench_c2_r$r.err
  python - <<PY
import json
d=[json.loads(l) for l in open("$OUT/bench_c2_r$r.json") if l.startswith("{")][-1]
e=d["e2e"]; print("run", $r, d["value"], e["value"], e["seconds"], e["upload_blocks"], e["phases_s"], [x["plan_s"] for x in e["blocks"]][:6])
PY
done
timeout 900 python -m pytest tests/test_gpu_mcmc_regression.py tests/test_gpu_stream_store.py -x -q 2>&1 | tail -3
