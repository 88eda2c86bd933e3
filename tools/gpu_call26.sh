#!/bin/bash
for v in "" abl/libdfgpu_arr2.so abl/libdfgpu_arr4.so abl/libdfgpu_ft3.so abl/libdfgpu_ft6.so; do
  echo "== lib: ${v:-default}"
  DFGPU_LIB=$v timeout 300 python tools/prof_l7.py --records 625000 2>&1 | grep -E "pack |flow_tag|total"
done
