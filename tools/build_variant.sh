#!/bin/bash
# Another build of the kernels beside the shipped library: tools/build_variant.sh TAG [-DFLAG ...] -> asmvar_b200/libage_b200_TAG.so
# (A/B measurements through AGE_LIB; e.g. `tools/build_variant.sh prof -DAGE_P2_PROF` for tools/gpu_prof.sh).  Needs the object
# files of the host side that __graft_entry__.build() leaves in asmvar_b200/csrc.
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root/asmvar_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function,-pthread \
     -I ../../include -I . "$@" -c age_kernels.cu -o /tmp/age_kernels_$tag.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a /tmp/age_kernels_$tag.o age_engine.o age_format.o -o ../libage_b200_$tag.so -lcudart
echo "built asmvar_b200/libage_b200_$tag.so"
