#!/bin/bash
# Round-2 GPU session C: parity after the state-layout refactor, second variant sweep, ncu.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02c_pytest.log 2>&1
grep -a "passed\|failed" $O/r02c_pytest.log | tail -2
V=pbrt-v1_b200/csrc/variants
{
for v in cur fuse fuse6 surf5 surf8 dir6c8; do
  echo -n "$v c3: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c3 2>&1 | tail -1
done
for v in cur fuse ds3 ds4 ds6 sh9; do
  for s in c4 c5; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s 2>&1 | tail -1
  done
done
} > $O/r02c_sweep.log 2>&1
cat $O/r02c_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path_surface|k_path_direct|k_path_continue" -s 3 -c 3 \
  -o $O/r02c_shade_d1 -f python tools/profile_render.py c3 > $O/r02c_ncu_shade.log 2>&1
tail -1 $O/r02c_ncu_shade.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -s 2 -c 2 \
  -o $O/r02c_trav_c5_d1 -f python tools/profile_render.py c5 --res 960 540 --spp 4 4 > $O/r02c_ncu_trav.log 2>&1
tail -1 $O/r02c_ncu_trav.log
