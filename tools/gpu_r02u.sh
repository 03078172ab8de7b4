#!/bin/bash
# Round-2 GPU session U: instance-capable walk with the hashed table inside object trees and 8-byte todo entries.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -av "^Rendering" ) > $O/r02u_pytest.log 2>&1
grep -a "passed\|failed\|Error" $O/r02u_pytest.log | tail -4
bash tools/inst_check.sh def gi_b5 gi_b7 gi_sb10 gi_sb6 > $O/r02u_inst.log 2>&1; grep "==\|full frame\|parity\|generate " $O/r02u_inst.log
