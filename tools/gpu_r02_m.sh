#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02p
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_rj.py tests/test_gpu_stream_store.py tests/test_gpu_mh.py -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --workload c3 --steps 60 --warmup 5 --no-cpu --no-extras > $OUT/bench_c3.json 2> $OUT/bench_c3.err
timeout 600 python bench.py --workload c5 --steps 60 --warmup 5 --no-cpu --no-extras > $OUT/bench_c5.json 2> $OUT/bench_c5.err
timeout 600 python bench.py --workload c5full --steps 60 --warmup 5 --no-cpu --no-extras > $OUT/bench_c5full.json 2> $OUT/bench_c5full.err
tail -n 4 $OUT/pytest.log
python - <<'PY'
import json
for w in ("c3","c5","c5full"):
    try:
        d=json.load(open(f"gpurun_out/r02p/bench_{w}.json")); print(w, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["phases_s"])
    except Exception as e: print(w, "failed", e)
PY
