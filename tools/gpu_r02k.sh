#!/bin/bash
# Round-2 GPU session K: GPU tests with the two-stream wave as default; stream modes and wave sizes A/B.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02k_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02k_pytest.log | tail -4
{
for s in c3 c4 c5; do
  echo "def $s: "; timeout 400 python tools/perf_sweep.py $s -- 5=0 "" 5=3 0=33554432 0=67108864 0=8388608 2=10 2=14 2=16 2>&1 | tail -9
done
} > $O/r02k_sweep.log 2>&1
cat $O/r02k_sweep.log
du -sh $O
