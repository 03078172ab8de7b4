#!/bin/bash
# Round-2 GPU session E: GPU tests on the parked k_path_fast build, sweep, bench line, launch list,
# ncu evidence for the TMA-staged-nodes variants (PB_SMEM_NODES) next to the default on C3 / C4 / C5.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/r02e_gpu.txt
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02e_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02e_pytest.log | tail -4
V=pbrt-v1_b200/csrc/variants
{
for v in cur smem256 smem2048; do
  for s in c3 c4 c5; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s 2>&1 | tail -1
  done
done
} > $O/r02e_sweep.log 2>&1
cat $O/r02e_sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02e_bench.json 2> $O/r02e_bench.err
tail -c 1500 $O/r02e_bench.json; tail -5 $O/r02e_bench.err
# launch list of the bench command (per-launch durations, cold cache / serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv --log-file $O/r02e_launches.csv \
  env PBRT_BENCH_LEAN=1 python bench.py --steps 2 --warmup 1 > $O/r02e_launches_bench.log 2>&1
tail -2 $O/r02e_launches_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path_fast" -s 1 -c 2 \
  -o $O/r02e_fast_d1 -f python tools/profile_render.py c3 > $O/r02e_ncu_fast.log 2>&1
tail -1 $O/r02e_ncu_fast.log
for v in cur smem2048; do
  for s in c4 c5; do
    extra=""; [ $s = c5 ] && extra="--res 960 540 --spp 4 4"
    [ $s = c4 ] && extra="--res 512 512 --spp 4 4"
    PBRT_B200_LIB=$PWD/$V/$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -s 2 -c 2 \
      -o $O/r02e_trav_${v}_${s} -f python tools/profile_render.py $s $extra > $O/r02e_ncu_trav_${v}_${s}.log 2>&1
    tail -1 $O/r02e_ncu_trav_${v}_${s}.log
  done
done
