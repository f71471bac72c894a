#!/bin/bash
# build a variant of the library with extra -D flags for gemm_tc.cu:  tools/build_variant.sh NAME -DSB_TC_SKIP=0 ...
# -> safepy_b200/libsafe_b200_NAME.so (select it with SAFE_B200_LIB); the other objects are those of the last build
set -e
name=$1; shift
cd "$(dirname "$0")/../safepy_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function --expt-relaxed-constexpr "$@" -c gemm_tc.cu -o /tmp/gemm_tc_$name.o
nvcc -shared -o ../libsafe_b200_$name.so neigh.o enrich.o /tmp/gemm_tc_$name.o graph.o finalize.o permstream.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built ../libsafe_b200_$name.so
