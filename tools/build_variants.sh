#!/bin/bash
# Scratch: build libsma_b200 variants with different conv_tc.cu compile-time knobs into variants_tmp/<name>.so
# usage: tools/build_variants.sh name1 "-DSMA_V2_UNROLL=6 -DSMA_EPI_PREFETCH=0" name2 "..." ...
set -e
cd "$(dirname "$0")/.."
CS=synergize-motion-appearance_b200/csrc
mkdir -p variants_tmp /tmp/var
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c $CS/conv_tc.cu -o /tmp/var/conv_tc_$name.o
  objs=$(ls $CS/build/*.o | grep -v conv_tc.o)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants_tmp/$name.so /tmp/var/conv_tc_$name.o $objs
  echo built $name
done
