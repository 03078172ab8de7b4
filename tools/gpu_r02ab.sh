#!/bin/bash
# Round-2 GPU session AB: wider differential fuzz of the GPU against the twin (every plugin mixed, instancing
# included) on the final walks: 4000 seeded scenes instead of the suite's 32.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( time PBRT_FUZZ_SCENES=4000 python -m pytest tests/test_gpu.py -m gpu -q -k "fuzz" 2>&1 | grep -av "^Rendering" ) > $O/r02ab_fuzz.log 2>&1
tail -6 $O/r02ab_fuzz.log
