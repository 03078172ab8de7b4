#!/bin/bash
# Builds a set of experimental variants of libpbrt_b200.so in parallel:
#   tools/build_variants.sh name1:"-DPB_X=1 -DPB_Y=2" name2:"" ...
# A name of the form rev@<git-rev> builds that revision's sources instead of the working tree.
cd "$(dirname "$0")/.."
ROOT=$PWD
mkdir -p pbrt-v1_b200/csrc/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Xptxas -v"
pids=()
for spec in "$@"; do
  name=${spec%%:*}; defs=""; [[ "$spec" == *:* ]] && defs=${spec#*:}
  src=$ROOT/pbrt-v1_b200/csrc
  if [[ "$name" == *@* ]]; then
    rev=${name#*@}; name=${name%%@*}
    tmp=$(mktemp -d)
    git archive "$rev" include pbrt-v1_b200/csrc | tar -x -C "$tmp"
    src=$tmp/pbrt-v1_b200/csrc
  fi
  ( cd "$src" && nvcc $FLAGS $defs -shared render.cu -o "$ROOT/pbrt-v1_b200/csrc/variants/$name.so" 2> "$ROOT/pbrt-v1_b200/csrc/variants/$name.ptxas.log" \
      && echo "built $name [$defs]" || echo "FAILED $name" ) &
  pids+=($!)
  if (( ${#pids[@]} >= 4 )); then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
done
wait
