#!/bin/bash
for v in "" abl/libdfgpu_kgseq.so; do
  echo "== lib: ${v:-default}"
  DFGPU_LIB=$v timeout 300 python tools/prof_l7.py --records 625000 2>&1 | grep -E "decode|smartencoding|total"
done
DFGPU_LIB=abl/libdfgpu_kgseq.so timeout 600 python -m pytest tests/test_l7_gpu.py -x -q 2>&1 | tail -2
