#!/bin/sh
# Build the C-ABI shared library in-tree for sm_100a.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -Xcompiler -fPIC -shared -o libfiat_b200.so fiat_b200.cu "$@"
