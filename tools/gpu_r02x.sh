#!/bin/bash
# Round-2 GPU session X (final tree): the tree as committed -- GPU tests, smoke, bench line (and the reference arm).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02x4_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02x4_pytest.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02x4_smoke.log 2>&1; tail -1 $O/r02x4_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02x4_bench.json 2> $O/r02x4_bench.err
tail -c 300 $O/r02x4_bench.json; tail -3 $O/r02x4_bench.err
