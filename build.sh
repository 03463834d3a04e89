#!/bin/bash
# Build libgpcounts_b200.so for sm_100a (also done by __graft_entry__.build()).
set -e
cd "$(dirname "$0")"
mkdir -p gpcounts_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" \
     -o gpcounts_b200/lib/libgpcounts_b200.so gpcounts_b200/csrc/gpc_api.cu gpcounts_b200/csrc/gpc_dense.cu
