#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFOO=1 ...]   ->  sid_b200/variants/libsidgpu_NAME.so (experiments only)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p sid_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" sid_b200/csrc/sidgpu.cu -o sid_b200/variants/libsidgpu_$name.so
