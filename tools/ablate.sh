#!/bin/bash
# builds ablation variants of l7_pass.cu (NOT parity-correct; for locating time only)
set -e
cd "$(dirname "$0")/.."
D=deepflow_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -x cu -I include"
for v in NO_KG NO_STRCOPY NO_ARRAYS; do
  nvcc $FLAGS -DABL_$v -c $D/l7_pass.cu -o /tmp/l7_pass_$v.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o abl/libdfgpu_$v.so $D/build/api.cu.o $D/build/dict_build.cu.o $D/build/framing.cu.o $D/build/l4.cu.o $D/build/l7.cu.o /tmp/l7_pass_$v.o -lcudart
done
