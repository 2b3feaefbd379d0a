#!/bin/sh
# Build the C-ABI shared library in-tree for sm_100a.  The translation units compile in parallel.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p _obj
pids=""
for tu in fiat_b200 small_launch vals_launch; do
    $NVCC $FLAGS "$@" -c -o _obj/$tu.o $tu.cu &
    pids="$pids $!"
done
for pid in $pids; do wait $pid; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libfiat_b200.so _obj/fiat_b200.o _obj/small_launch.o _obj/vals_launch.o
