#!/bin/bash
# Builds an experimental variant of the CUDA library with extra -D flags:
#   tools/build_variant.sh NAME -DPB_X=1 ...   ->  pbrt-v1_b200/csrc/variants/NAME.so
# Select it at run time with PBRT_B200_LIB=pbrt-v1_b200/csrc/variants/NAME.so (pyffi.py).
set -e
cd "$(dirname "$0")/../pbrt-v1_b200/csrc"
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
  -ftz=false -Xcompiler -fPIC -Xptxas -v "$@" -shared render.cu -o variants/$name.so 2> variants/$name.ptxas.log
grep -A2 "k_intersectILb0\|k_shadowILb0" variants/$name.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo
