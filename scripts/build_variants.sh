#!/bin/bash
# build tuning variants of the library: libfa_v_<name>.so with extra -D flags:  build_variants.sh name1 "-DX=1" name2 "-DY=2" ...
cd "$(dirname "$0")/.."; mkdir -p variants
while [ $# -gt 1 ]; do
  nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 $2 -shared -Xcompiler -fPIC -o variants/libfa_v_$1.so flash-attention-cuda-c_b200/kernels/FlashAttention.cu || exit 1
  shift 2
done
ls variants
