#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02v
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/pytest.log
for w in c2 c1 c4a c4b; do
timeout 600 python bench.py --workload $w --no-cpu --no-extras > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"
done
python - <<'PY'
import json
for w in ("c2","c1","c4a","c4b"):
    d=[json.loads(l) for l in open(f"gpurun_out/r02v/bench_{w}.json") if l.startswith("{")][-1]
    print(w, d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"] and d["e2e"]["value"], d["roofline"].get("frac"), d["roofline"].get("kernel_ms"))
PY
