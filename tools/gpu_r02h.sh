#!/bin/bash
# Round-2 GPU session H: tuning sweep of the lean walk (refill threshold, interior-phase bar, leaf steps, occupancy,
# mailbox slots) on C3/C4 (+C5 for the front runners); ncu summary of C3's traversal kernels.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02h_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02h_pytest.log | tail -4
V=pbrt-v1_b200/csrc/variants
{
for v in def rf8 rf24 lanes8 lanes16 fix4 max12 t10 t8 sh9 sh12 ls1 ls3 mb32; do
  for s in c3 c4; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s 2>&1 | tail -1
  done
done
for v in def rf8 lanes8 sh12 t8; do
  echo -n "$v c5: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c5 2>&1 | tail -1
done
} > $O/r02h_sweep.log 2>&1
cat $O/r02h_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -s 2 -c 2 \
  -o /tmp/trav_c3 -f python tools/profile_render.py c3 > $O/r02h_ncu_trav_c3.log 2>&1
python tools/ncu_summary.py /tmp/trav_c3.ncu-rep 40 > $O/r02h_ncu_trav_def_c3.txt 2>&1
tail -1 $O/r02h_ncu_trav_c3.log
du -sh $O
