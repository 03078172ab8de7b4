#!/bin/bash
# Round-2 GPU session AA: persisting-L2 window over the kd-tree nodes (tunable 11) on C4 / C5.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
{
for s in c4 c5; do
  echo "def $s: "; timeout 400 python tools/perf_sweep.py $s -- "" 11=25 11=50 11=75 2>&1 | tail -4
done
} > $O/r02aa_sweep.log 2>&1
cat $O/r02aa_sweep.log
