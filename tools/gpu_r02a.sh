#!/bin/bash
# Round-2 GPU session A: parity (incl. the at-scale tests against the live reference), variant sweep, ncu.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/r02a_gpu.txt; nproc >> $O/r02a_gpu.txt
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^Rendering" ) > $O/r02a_pytest.log 2>&1
tail -3 $O/r02a_pytest.log
V=pbrt-v1_b200/csrc/variants
{
for v in r1 cur nolam mb4 mb8 occ1 occ2; do
  for s in c3 c4 c5; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s 2>&1 | tail -1
  done
done
echo -n "cur(noframes) c3: "; PBRT_B200_LIB=$PWD/$V/cur.so timeout 300 python tools/perf_sweep.py c3 -- 7=0 2>&1 | tail -1
echo -n "cur(noframes) c4: "; PBRT_B200_LIB=$PWD/$V/cur.so timeout 300 python tools/perf_sweep.py c4 -- 7=0 2>&1 | tail -1
} > $O/r02a_sweep.log 2>&1
cat $O/r02a_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path_surface|k_path_direct|k_path_continue" -s 3 -c 3 \
  -o $O/r02a_shade_d1 -f python tools/profile_render.py c3 > $O/r02a_ncu_shade.log 2>&1
tail -2 $O/r02a_ncu_shade.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02a_bench.json 2> $O/r02a_bench.err
cat $O/r02a_bench.json | cut -c1-600
