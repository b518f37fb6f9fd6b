#!/usr/bin/env bash
# Builds libbd_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr -cudart static"
$NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c bd_api.cu -o bd_api.o
$NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c post.cu -o post.o
$NVCC $FLAGS -fmad=false ${PTXAS_V:+-Xptxas -v} -c contours.cu -o contours.o
$NVCC $FLAGS -c png0.cpp -o png0.o
$NVCC $FLAGS -shared bd_api.o post.o contours.o png0.o -o ../libbd_b200.so
echo "built $(cd .. && pwd)/libbd_b200.so"
