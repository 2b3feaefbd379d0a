#!/bin/sh
# Build the C-ABI shared library in-tree for sm_100a.  The translation units compile in parallel;
# an object is rebuilt only if one of the sources it includes is newer (FORCE=1 rebuilds all).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p _obj
deps_fiat_b200="fiat_b200.cu host_plan.cuh device_plan.cuh expansion.cuh kernels.cuh lattice.cuh ../../include/fiat_b200.h"
deps_small_launch="small_launch.cu host_plan.cuh device_plan.cuh expansion.cuh small.cuh ../../include/fiat_b200.h"
deps_cells_launch="cells_launch.cu host_plan.cuh device_plan.cuh expansion.cuh cells.cuh cells_reg.cuh ../../include/fiat_b200.h"
deps_vals_launch="vals_launch.cu host_plan.cuh device_plan.cuh expansion.cuh small.cuh vals.cuh ../../include/fiat_b200.h"
deps_cluster="cluster.cu ../../include/fiat_b200.h"
pids=""
for tu in fiat_b200 small_launch vals_launch cells_launch cluster; do
    eval deps=\$deps_$tu
    stale=0
    [ -n "$FORCE" ] || [ -n "$*" ] || [ ! -f _obj/$tu.o ] && stale=1
    for f in $deps build.sh; do [ "$f" -nt _obj/$tu.o ] && stale=1; done
    if [ $stale = 1 ]; then
        $NVCC $FLAGS "$@" -c -o _obj/$tu.o $tu.cu &
        pids="$pids $!"
    fi
done
for pid in $pids; do wait $pid; done
# (linked under a temporary name and renamed: a snapshot taken meanwhile never sees a half-written library)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o _obj/libfiat_b200.so.tmp _obj/fiat_b200.o _obj/small_launch.o _obj/vals_launch.o _obj/cells_launch.o _obj/cluster.o
mv -f _obj/libfiat_b200.so.tmp libfiat_b200.so
