#!/bin/bash
# builds tuning variants of the L7 kernels into abl/libdfgpu_<name>.so (select with DFGPU_LIB=...)
#   tools/variants.sh name file.cu "-DFLAG=.. -DFLAG2=.."
set -e
cd "$(dirname "$0")/.."
D=deepflow_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -x cu -I include"
name=$1; file=$2; defs=$3
mkdir -p abl
nvcc $FLAGS $defs -c $D/$file -o /tmp/var_$name.o
objs=""
for o in $D/build/*.o; do
  case "$o" in *"/$file.o") objs="$objs /tmp/var_$name.o";; *) objs="$objs $o";; esac
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o abl/libdfgpu_$name.so $objs -lcudart
echo abl/libdfgpu_$name.so
