#!/bin/bash
# Round-2 GPU session I: parked quadric tests (PB_QUADRIC_LANES), size-dependent refill threshold, interior-phase bar.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02i_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02i_pytest.log | tail -4
V=pbrt-v1_b200/csrc/variants
{
for v in def q1 q2 q8 lanes6 lanes4 rfs12 rfs20 q4t10 q4t12; do
  echo -n "$v c3: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c3 2>&1 | tail -1
done
for v in def rfb4 rfb12 lanes6 lanes4; do
  echo -n "$v c4: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c4 2>&1 | tail -1
done
for v in def rfb4 lanes6; do
  echo -n "$v c5: "; PBRT_B200_LIB=$PWD/$V/$v.so timeout 300 python tools/perf_sweep.py c5 2>&1 | tail -1
done
} > $O/r02i_sweep.log 2>&1
cat $O/r02i_sweep.log
PBRT_PROFILE_COUNTS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -c 6 \
  -o /tmp/trav_c3 -f python tools/profile_render.py c3 > $O/r02i_ncu_trav_c3.log 2>&1
python tools/ncu_summary.py /tmp/trav_c3.ncu-rep 40 > $O/r02i_ncu_trav_def_c3.txt 2>&1
python tools/traffic_from_ncu.py /tmp/trav_c3.ncu-rep $O/r02i_ncu_trav_c3.log "profiles/r02i_ncu_trav_def_c3.txt (ncu --set full, first 3 closest-hit + 3 any-hit launches of wave 1, C3)" > $O/r02i_traffic.json 2>$O/r02i_traffic.err
cat $O/r02i_traffic.json
du -sh $O
