#!/bin/bash
# Round-2 GPU session G: treelet node order (tunable 9), 8-byte todo entries, inline-pop on/off, mailbox table per ray kind.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02g_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02g_pytest.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02g_smoke.log 2>&1; tail -2 $O/r02g_smoke.log
V=pbrt-v1_b200/csrc/variants
{
for v in cur def nopop mb0 mb0_nopop mb16_any0 mb16_any0_nopop mb8_any0_nopop mb16_any4_nopop; do
  for s in c3 c4; do
    echo "$v $s: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py $s -- "" 9=0 2>&1 | tail -2
  done
done
for v in def nopop mb16_any0_nopop mb0_nopop; do
  echo "$v c5: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c5 -- "" 9=0 2>&1 | tail -2
done
} > $O/r02g_sweep.log 2>&1
cat $O/r02g_sweep.log
for v in def; do
  for s in c4 c5; do
    extra="--res 512 512 --spp 4 4"; [ $s = c5 ] && extra="--res 960 540 --spp 4 4"
    PBRT_B200_LIB=$PWD/$V/$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -s 2 -c 2 \
      -o /tmp/trav_${v}_${s} -f python tools/profile_render.py $s $extra > $O/r02g_ncu_trav_${v}_${s}.log 2>&1
    python tools/ncu_summary.py /tmp/trav_${v}_${s}.ncu-rep 30 > $O/r02g_ncu_trav_${v}_${s}.txt 2>&1
    tail -1 $O/r02g_ncu_trav_${v}_${s}.log
  done
done
du -sh $O
