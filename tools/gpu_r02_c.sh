#!/bin/bash
# round 2: whole GPU suite + the new bench line (c2 default with c3 sub-record) + reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02c
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" | tee -a $OUT/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench c2 rc=$?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "bench ref rc=$?" | tee -a $OUT/summary.txt
tail -n 5 $OUT/pytest_gpu.log; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c/bench_c2.json"))
for k in ("value","ms_per_step","e2e","cpu_baseline","sweep_forms","with_fitted_values","ess_long","c3"):
    print(k, json.dumps(d.get(k))[:900])
print("roofline", json.dumps(d["roofline"])[:1200])
PY
tail -n 5 $OUT/bench_c2.err; cat $OUT/bench_ref.json | head -c 600
