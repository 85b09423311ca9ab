#!/bin/sh
# tools/build_variant.sh <name> <nvcc defines...>: pcdet_b200/libpcdet_b200_<name>.so with sparse_conv_tc.cu compiled with the defines
set -e
cd "$(dirname "$0")/../pcdet_b200"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/sparse_conv_tc.cu -o /tmp/sparse_conv_tc_$name.o
OBJS=$(ls csrc/_obj/*.o | grep -v sparse_conv_tc.o)
nvcc -shared -o libpcdet_b200_$name.so $OBJS /tmp/sparse_conv_tc_$name.o -gencode arch=compute_100a,code=sm_100a
echo built libpcdet_b200_$name.so
