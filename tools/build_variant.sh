#!/bin/bash
# usage: build_variant.sh NAME [-DPA_...=x ...]   -> variants/libv_NAME.so (attention_pair.cu recompiled with the flags,
# every other object taken from the product build).  Bring-up tool for A/B runs: AMB_PROBE_LIB=variants/libv_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" \
  -c actionmesh_b200/csrc/attention_pair.cu -o variants/attention_pair_$name.o
objs=$(ls actionmesh_b200/lib/*.o | grep -v attention_pair.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libv_$name.so $objs variants/attention_pair_$name.o -lcudart_static -ldl -lrt -lpthread
echo built variants/libv_$name.so
