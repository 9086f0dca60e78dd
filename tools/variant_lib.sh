#!/bin/bash
# Build a variant of libuof_b200.so with one source recompiled under extra defines (A/B kernels without touching the
# default build):  bash tools/variant_lib.sh <name> <source.cu> "<-D flags>"   ->  unopticalflow_b200/lib/variants/libuof_<name>.so
# Select it at run time with UOF_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../unopticalflow_b200"
name=$1; src=$2; flags=$3
mkdir -p lib/variants
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $flags -c csrc/$src -o lib/variants/${src%.cu}_$name.o
objs=$(ls lib/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o lib/variants/libuof_$name.so $objs lib/variants/${src%.cu}_$name.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo lib/variants/libuof_$name.so
