#!/bin/bash
# Round-2 GPU session N: two wave lanes (tunable 10): GPU tests, A/B on C3 / C4 / C5 (one lane, two lanes with
# half-size waves = same memory, two lanes with full-size waves = twice the memory), 8 Mi-sample frames like a
# multi-GPU rank's share.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02n_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02n_pytest.log | tail -4
{
for s in c3 c4 c5; do
  echo "def $s: "; timeout 400 python tools/perf_sweep.py $s -- 10=1 "" 0=134217728 2>&1 | tail -3
done
echo "c3 at 362x362 (8.4 M samples, an 8-GPU rank's share): "; timeout 300 python tools/perf_sweep.py c3 --res 362 362 -- 10=1 "" 2>&1 | tail -2
} > $O/r02n_sweep.log 2>&1
cat $O/r02n_sweep.log
du -sh $O
