#!/bin/bash
# usage: tools/inst_check.sh VARIANT...   instanced benchmark scene under variant libraries
cd "$(dirname "$0")/.."
for v in "$@"; do
  echo "== $v"
  PBRT_B200_LIB=$PWD/pbrt-v1_b200/csrc/variants/$v.so python tools/big_scene_check.py inst --full 2>&1 | grep -E "parity|full frame|generate "
done
