#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r02d
mkdir -p $OUT
timeout 600 python tools/profile_prepare.py c2 > $OUT/prof_c2.log 2>&1
timeout 600 python tools/profile_prepare.py c3 > $OUT/prof_c3.log 2>&1
timeout 900 python -m pytest tests/test_gpu_mh.py tests/test_gpu_stream_store.py tests/test_gpu_rj.py -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -n 4 $OUT/pytest.log
