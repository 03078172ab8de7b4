#!/bin/bash
# Round-2 GPU session V: k_path_fast variants (prefetch of the parked hits' records, occupancy).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=pbrt-v1_b200/csrc/variants
{
for v in def fpf fast5 fpf5; do
  for s in c3 c4; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s -- 5=0 2>&1 | tail -1
  done
done
} > $O/r02v_sweep.log 2>&1
cat $O/r02v_sweep.log
