#!/bin/bash
# Round-2 GPU session AG: 12000-seed GPU-vs-twin fuzz on the final tree.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( PBRT_FUZZ_SCENES=12000 python -m pytest tests/test_gpu.py -m gpu -q -k "fuzz" -p no:cacheprovider 2>&1 | grep -av "^Rendering" ) > $O/r02ag_fuzz.log 2>&1; grep -a "passed\|failed\|FAILED" $O/r02ag_fuzz.log | tail -8
