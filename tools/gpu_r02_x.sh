#!/bin/bash
# round 2, run X: full GPU suite + the default bench line + the reference arm (what the driver runs at round end)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02x
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
OMC_BENCH_E2E_PREP=stat timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$?"; grep "allocator state" $OUT/bench_default.err
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench reference rc=$?"
timeout 600 python bench.py --workload c3 --no-cpu --no-extras > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02x/bench_default.json") if l.startswith("{")][-1]
print("C2", d["value"], d["ms_per_step"], d["gpu_launches"], "e2e", d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["upload_blocks"])
print(" c3", d["c3"]["value"], d["c3"]["roofline"]["frac"], d["c3"]["e2e"]["value"], d["c3"]["e2e"]["phases_s"])
r=[json.loads(l) for l in open("gpurun_out/r02x/bench_reference.json") if l.startswith("{")][-1]
print("REF", r["value"], r["steps"], r["warmup"])
c=[json.loads(l) for l in open("gpurun_out/r02x/bench_c3.json") if l.startswith("{")][-1]
print("c3 alone", c["value"], c["e2e"]["value"], c["e2e"]["phases_s"])
PY
