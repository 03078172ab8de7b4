#!/bin/bash
# Round-2 GPU session L: 64 Mi-sample waves + three-stream wave as defaults: GPU tests, larger waves, k_path_fast
# occupancy / out-of-line mixer variants, full bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02l_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02l_pytest.log | tail -4
V=pbrt-v1_b200/csrc/variants
{
for s in c3 c4 c5; do
  echo "def $s: "; timeout 400 python tools/perf_sweep.py $s -- "" 0=134217728 0=268435456 5=1 2>&1 | tail -4
done
for v in fast7 fast8 mixcall; do
  echo -n "$v c3: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c3 2>&1 | tail -1
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
} > $O/r02l_sweep.log 2>&1
cat $O/r02l_sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02l_bench.json 2> $O/r02l_bench.err
tail -c 300 $O/r02l_bench.json; tail -3 $O/r02l_bench.err
du -sh $O
