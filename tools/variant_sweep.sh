#!/bin/bash
# usage: tools/variant_sweep.sh "c3 c4" v2 v3q16[:tunables] ...   (runs tools/perf_sweep.py per variant library)
cd "$(dirname "$0")/.."
scenes=$1; shift
for vv in "$@"; do
  v=${vv%%:*}; tn=""; [[ "$vv" == *:* ]] && tn="-- ${vv#*:}"
  for s in $scenes; do
    echo -n "$vv $s: "
    PBRT_B200_LIB=$PWD/pbrt-v1_b200/csrc/variants/$v.so python tools/perf_sweep.py $s $tn 2>&1 | tail -1
  done
done
