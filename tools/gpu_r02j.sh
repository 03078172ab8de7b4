#!/bin/bash
# Round-2 GPU session J: final walk defaults; probes on the aux stream (tunable 5 = 2) A/B; full bench line;
# launch list of the bench command; ncu summaries of the bounce launches of C3 / C4 / C5.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02j_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02j_pytest.log | tail -4
{
for s in c3 c4 c5; do
  echo "def $s: "; timeout 300 python tools/perf_sweep.py $s -- "" 5=2 5=1 2>&1 | tail -3
done
PBRT_PROFILE_COUNTS=1 timeout 300 python tools/profile_render.py c5 2>&1 | grep -a "iteration\|render" | head -8
} > $O/r02j_sweep.log 2>&1
cat $O/r02j_sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02j_bench.json 2> $O/r02j_bench.err
tail -c 600 $O/r02j_bench.json; tail -3 $O/r02j_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv --log-file $O/r02j_launches.csv \
  env PBRT_BENCH_LEAN=1 python bench.py --steps 2 --warmup 1 > $O/r02j_launches_bench.log 2>&1
tail -c 300 $O/r02j_launches_bench.log
for s in c3 c4 c5; do
  extra=""; [ $s = c4 ] && extra="--res 512 512 --spp 4 4"; [ $s = c5 ] && extra="--res 960 540 --spp 4 4"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -s 2 -c 2 \
    -o /tmp/trav_$s -f python tools/profile_render.py $s $extra > $O/r02j_ncu_trav_$s.log 2>&1
  python tools/ncu_summary.py /tmp/trav_$s.ncu-rep 30 > $O/r02j_ncu_trav_$s.txt 2>&1
  tail -1 $O/r02j_ncu_trav_$s.log
done
du -sh $O
