#!/bin/bash
# Round-2 GPU session D: tile mode / fast kernel parity, occupancy sweep of k_path_fast, new bench line, ncu.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02d_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02d_pytest.log | tail -4
V=pbrt-v1_b200/csrc/variants
{
for v in cur fast3 fast5 fast6; do
  for s in c3 c4 c5; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s 2>&1 | tail -1
  done
done
echo -n "cur(nofast) c3: "; PBRT_B200_LIB=$PWD/$V/cur.so timeout 300 python tools/perf_sweep.py c3 -- 8=0 2>&1 | tail -1
for v in ds4 ds3l3; do
  for s in c3 c4; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s 2>&1 | tail -1
  done
done
} > $O/r02d_sweep.log 2>&1
cat $O/r02d_sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02d_bench.json 2> $O/r02d_bench.err
tail -c 3000 $O/r02d_bench.json; tail -5 $O/r02d_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path_fast" -s 1 -c 2 \
  -o $O/r02d_fast_d1 -f python tools/profile_render.py c3 > $O/r02d_ncu_fast.log 2>&1
tail -1 $O/r02d_ncu_fast.log
