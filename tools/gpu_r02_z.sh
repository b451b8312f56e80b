#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02z
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 900 python bench.py --no-cpu > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02z/bench_default.json") if l.startswith("{")][-1]
print("C2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["upload_blocks"])
print(" fitted", d["with_fitted_values"])
print(" ess_long", d["ess_long"]["value"], d["ess_long"].get("bulk"))
print(" c3", d["c3"]["value"], d["c3"]["e2e"]["value"])
PY
