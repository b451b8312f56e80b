#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02y
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_rj.py -x -q > $OUT/pytest_rj.log 2>&1; echo "pytest rj rc=$?"; tail -3 $OUT/pytest_rj.log
for w in c5 c5full; do
  timeout 600 python bench.py --workload $w --no-cpu --no-extras > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - <<PY
import json
d=[json.loads(l) for l in open("$OUT/bench_$w.json") if l.startswith("{")][-1]
print("$w", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"])
PY
done
for prep in stat empty warm; do
  OMC_BENCH_E2E_PREP=$prep timeout 900 python bench.py --no-cpu > $OUT/bench_$prep.json 2> $OUT/bench_$prep.err
  grep "allocator state" $OUT/bench_$prep.err
  python - <<PY
import json
d=[json.loads(l) for l in open("$OUT/bench_$prep.json") if l.startswith("{")][-1]
e=d["e2e"]; print("$prep", e["value"], e["seconds"], e["upload_blocks"], e["phases_s"], [(b["queue_next_s"], b["plan_s"]) for b in e["blocks"]][:5])
PY
done
