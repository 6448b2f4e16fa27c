#!/bin/bash
# Build tuning variants of the library: scripts/build_variants.sh name "-DFLAG=.." [name2 "..."] -> build/variants/lib_<name>.so
mkdir -p build/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -diag-suppress 1444 \
    -Wno-deprecated-declarations $flags -I include -o build/variants/lib_$name.so allset_b200/csrc/allset_kernels.cu &
done
wait
ls -la build/variants
