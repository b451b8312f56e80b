#!/bin/bash
# round 2: timing of the tridiagonal draw (shipped library + any build variants in tools/variants) + the gmrf tests
mkdir -p gpurun_out/r02q
{ echo "== shipped"; python tools/perf_tridiag.py 2>&1 | tail -1; python tools/perf_tridiag.py 2>&1 | tail -1
for f in tools/variants/libomc_td_*.so; do
  [ -f "$f" ] || continue
  echo "== $f"; OMC_LIB=$f python tools/perf_tridiag.py 2>&1 | tail -1
done; } > gpurun_out/r02q/variants.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_gmrf.py tests/test_gpu_gmrf_module.py tests/test_gpu_stream_store.py -x -q > gpurun_out/r02q/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02q/variants.txt
tail -3 gpurun_out/r02q/pytest.log >> gpurun_out/r02q/variants.txt
cat gpurun_out/r02q/variants.txt
