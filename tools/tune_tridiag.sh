#!/bin/bash
# Build tridiag variants (tuning knobs) into tools/variants/ and time each on the GPU (run the timing part under gpurun).
cd "$(dirname "$0")/.."
OUT=tools/variants; mkdir -p $OUT
SRC=openmcmc_b200/csrc
build() { # name flags...   (only tridiag.cu is recompiled; the other objects come from the regular build)
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
     -c $SRC/tridiag.cu -o $OUT/$name.o 2> $OUT/$name.log || { tail -5 $OUT/$name.log; return 1; }
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libomc_$name.so $OUT/$name.o \
     $(ls $SRC/build/*.o | grep -v tridiag.o) && rm -f $OUT/$name.o
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
