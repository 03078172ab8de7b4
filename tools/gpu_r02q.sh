#!/bin/bash
# Round-2 GPU session Q: instance-capable walk with the adaptive interior phase as default: GPU tests; tuning variants
# on the 27-placement scene; bench line (e2e leg with W warm-up steps).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02q_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02q_pytest.log | tail -4
bash tools/inst_check.sh def gi_l6 gi_l10 gi_r4 gi_r12 gi_b8 gi_b5 gi_max12 > $O/r02q_inst.log 2>&1; grep "==\|full frame" $O/r02q_inst.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02q_bench.json 2> $O/r02q_bench.err
tail -c 300 $O/r02q_bench.json; tail -3 $O/r02q_bench.err
