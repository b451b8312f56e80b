#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02w
mkdir -p $OUT
for b in 5 8 10 16; do
  timeout 600 python bench.py --workload c2 --no-cpu --no-extras --upload-blocks $b > $OUT/bench_c2_b$b.json 2> $OUT/bench_c2_b$b.err
  python - <<PY
import json
d=[json.loads(l) for l in open("$OUT/bench_c2_b$b.json") if l.startswith("{")][-1]
e=d["e2e"]; print("blocks", $b, e["value"], e["seconds"], e["phases_s"], [x["plan_s"] for x in e["blocks"]][:4])
PY
done
for ex in 1_model_distributions 2_samplers 3_linear_regression; do
  timeout 600 python examples/$ex.py > $OUT/example_$ex.log 2>&1; echo "example $ex rc=$?"; tail -2 $OUT/example_$ex.log
done
timeout 600 python examples/4_GMRF_smoother.py 1000000 4 > $OUT/example_4.log 2>&1; echo "example 4 rc=$?"; tail -3 $OUT/example_4.log
