#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02k
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gmrf.py tests/test_gpu_gmrf_module.py tests/test_gpu_mcmc_regression.py tests/test_gpu_fused_small.py tests/test_gpu_rj.py -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --workload c3 --steps 60 --warmup 5 --no-cpu --no-extras > $OUT/bench_c3.json 2> $OUT/bench_c3.err
echo "bench c3 rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none -k regex:tg_ -s 6 -c 4 -o $OUT/ncu_c3 python bench.py --workload c3 --steps 10 --warmup 2 --no-cpu --no-extras --no-e2e > $OUT/ncu.log 2>&1
tail -n 4 $OUT/pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02k/bench_c3.json")); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], json.dumps(d["e2e"])[:600])
PY
