#!/bin/bash
# round 2, run U: N GPUs of one box -- the bench line the driver launches (torchrun, one rank per GPU), C2 and C3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
OUT=gpurun_out/r02u$N
mkdir -p $OUT
nvidia-smi -L | head -8
for w in c2 c3; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  echo "bench $w N=$N rc=$?"; tail -c 600 $OUT/bench_$w.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "reference rc=$?"
python - <<PY
import json
for w in ("c2","c3"):
    try:
        d=[json.loads(l) for l in open("$OUT/bench_%s.json"%w) if l.startswith("{")][-1]
        print(w, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"].get("pinned_policy")), "strong", d.get("strong"))
    except Exception as e: print(w, "failed", e)
PY
