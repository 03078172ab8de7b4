#!/bin/bash
# Round-2 GPU session M (final state): GPU tests, smoke, the bench line, the ncu launch list of the bench command,
# ncu summaries of the traversal kernels (bounce launches) and of k_path_fast, refreshed traffic.json.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02m_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02m_pytest.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02m_smoke.log 2>&1; tail -1 $O/r02m_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02m_bench.json 2> $O/r02m_bench.err
tail -c 300 $O/r02m_bench.json; tail -3 $O/r02m_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02m_bench_reference.json 2> $O/r02m_bench_reference.err
tail -c 300 $O/r02m_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/r02m_launches.csv \
  env PBRT_BENCH_LEAN=1 python bench.py --steps 2 --warmup 1 > $O/r02m_launches_bench.log 2>&1
tail -c 200 $O/r02m_launches_bench.log
export PBRT_PROFILE_WAVE=16777216
for s in c3 c4 c5; do
  extra=""; [ $s = c4 ] && extra="--res 512 512 --spp 4 4"; [ $s = c5 ] && extra="--res 960 540 --spp 4 4"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shadow" -s 2 -c 2 \
    -o /tmp/trav_$s -f python tools/profile_render.py $s $extra > $O/r02m_ncu_trav_$s.log 2>&1
  python tools/ncu_summary.py /tmp/trav_$s.ncu-rep 30 > $O/r02m_ncu_trav_$s.txt 2>&1
  tail -1 $O/r02m_ncu_trav_$s.log
done
PBRT_PROFILE_COUNTS=1 timeout 600 ncu --set full --clock-control none -k regex:"k_intersect|k_shadow" -c 6 \
  -o /tmp/traffic_c3 -f python tools/profile_render.py c3 > $O/r02m_ncu_traffic_c3.log 2>&1
python tools/traffic_from_ncu.py /tmp/traffic_c3.ncu-rep $O/r02m_ncu_traffic_c3.log "ncu --set full, first 3 closest-hit + 3 any-hit launches of the frame's wave, C3 (tools/gpu_r02m.sh)" > $O/r02m_traffic.json 2>$O/r02m_traffic.err
cat $O/r02m_traffic.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path_fast" -s 1 -c 1 \
  -o /tmp/fast_d1 -f python tools/profile_render.py c3 > $O/r02m_ncu_fast.log 2>&1
python tools/ncu_summary.py /tmp/fast_d1.ncu-rep 30 > $O/r02m_ncu_k_path_fast.txt 2>&1
du -sh $O
