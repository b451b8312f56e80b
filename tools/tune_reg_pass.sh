#!/bin/bash
# Build reg_pass variants (tuning knobs) into gpurun_out/variants/ and time each on the GPU (run under gpurun).
set -e
cd "$(dirname "$0")/.."
OUT=tools/variants; mkdir -p $OUT
SRC=openmcmc_b200/csrc
build() { # name flags...
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
     $SRC/reg_pass.cu $SRC/omc_api.cu $SRC/dense_draw.cu $SRC/logp.cu -o $OUT/libomc_$name.so 2> $OUT/$name.log || { tail -5 $OUT/$name.log; return 1; }
}
if [ "$1" == "build" ]; then
  build tg2_bulk &
  build tg1_bulk -DOMC_RP_TG=1 &
  build tg2_ldgsts -DOMC_RP_USE_BULK=0 &
  build tg1_bulk_kc64 -DOMC_RP_TG=1 -DOMC_RP_KC=64 -DOMC_RP_NSTAGE=3 &
  wait
  build tg2_bulk_kc64s2 -DOMC_RP_KC=64 -DOMC_RP_NSTAGE=2 &
  build tg2_bulk_kc16s8 -DOMC_RP_KC=16 -DOMC_RP_NSTAGE=8 &
  wait
  ls -la $OUT
else
  for f in $OUT/libomc_*.so; do
    echo "== $f"; OMC_LIB=$f python tools/perf_reg_pass.py ${1:-2048} 2>&1 | tail -1
  done
fi
