#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02i
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
for w in c2 c1 c4a c4b; do
  timeout 600 python bench.py --workload $w --steps 60 --warmup 5 --no-cpu --no-extras --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  echo "bench $w rc=$?" | tee -a $OUT/summary.txt
done
tail -n 4 $OUT/pytest.log
python - <<'PY'
import json
for w in ("c2","c1","c4a","c4b"):
    try:
        d=json.load(open(f"gpurun_out/r02i/bench_{w}.json")); print(w, d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["kernel_ms"])
    except Exception as e: print(w, "failed", e)
PY
