#!/bin/bash
# Builds libflowchain_b200.so (C-ABI, include/flowchain_b200.h) in-tree for sm_100a.
set -e
cd "$(dirname "$0")"
OUT=../libflowchain_b200.so
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
     -Xptxas -v -o "$OUT" capi.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning: v|spill" build.log | grep -v "0 bytes spill" | head -20 || true
echo "built $OUT"
