#!/bin/bash
# Build tridiag variants (tuning knobs) into tools/variants/ and time each on the GPU (run the timing part under gpurun).
cd "$(dirname "$0")/.."
OUT=tools/variants; mkdir -p $OUT
SRC=openmcmc_b200/csrc
build() { # name flags...
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v "$@" \
     $SRC/tridiag.cu $SRC/omc_api.cu $SRC/dense_draw.cu $SRC/logp.cu $SRC/reg_pass.cu $SRC/mh.cu -o $OUT/libomc_$name.so 2> $OUT/$name.log || { tail -5 $OUT/$name.log; return 1; }
}
if [ "$1" == "build" ]; then
  shift
  rm -f $OUT/libomc_td_*.so
  i=0
  for v in "$@"; do
    name=td_$(echo "$v" | tr -d ' =' | tr -c 'A-Za-z0-9_\n' '_')
    build $name $v &
    i=$((i+1)); if [ $((i % 4)) == 0 ]; then wait; fi
  done
  wait
  ls $OUT | grep td_
else
  for f in $OUT/libomc_td_*.so; do
    echo "== $f"; OMC_LIB=$f python tools/perf_tridiag.py 2>&1 | tail -1
  done
fi
