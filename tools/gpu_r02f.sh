#!/bin/bash
# Round-2 GPU session F: lean kd walk (hashed mailbox table, inline empty-leaf pop, adaptive interior phase):
# GPU tests, variant sweep on C3/C4/C5, ncu summaries (reports are summarised on the box: they are too big to travel).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02f_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02f_pytest.log | tail -4
V=pbrt-v1_b200/csrc/variants
{
for v in cur lean0 lean8 lean16 l16_lanes8 l16_lanes16 l16_ls1 l16_ls3 l16_t10 l16_fix4 l16_sh9 l16_p8; do
  for s in c3 c4; do
    echo -n "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s 2>&1 | tail -1
  done
done
for v in lean0 lean16 l16_t10 l16_p8 l16_lanes8; do
  echo -n "$v c5: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c5 2>&1 | tail -1
done
} > $O/r02f_sweep.log 2>&1
cat $O/r02f_sweep.log
for v in cur smem2048 lean16; do
  for s in c4 c5; do
    extra="--res 512 512 --spp 4 4"; [ $s = c5 ] && extra="--res 960 540 --spp 4 4"
    PBRT_B200_LIB=$PWD/$V/$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -s 2 -c 2 \
      -o /tmp/trav_${v}_${s} -f python tools/profile_render.py $s $extra > $O/r02f_ncu_trav_${v}_${s}.log 2>&1
    python tools/ncu_summary.py /tmp/trav_${v}_${s}.ncu-rep 30 > $O/r02f_ncu_trav_${v}_${s}.txt 2>&1
    tail -1 $O/r02f_ncu_trav_${v}_${s}.log
  done
done
cp /tmp/trav_lean16_c5.ncu-rep $O/r02f_trav_lean16_c5.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path_fast" -s 1 -c 2 \
  -o /tmp/fast_d1 -f python tools/profile_render.py c3 > $O/r02f_ncu_fast.log 2>&1
python tools/ncu_summary.py /tmp/fast_d1.ncu-rep 30 > $O/r02f_ncu_fast.txt 2>&1
cp /tmp/fast_d1.ncu-rep $O/r02f_fast_d1.ncu-rep
du -sh $O
