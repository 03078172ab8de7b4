#!/bin/bash
# Round-2 GPU session AD: x86-faithful float-to-int conversion on the device: the two fuzz edge cases, the 4000-seed
# fuzz again, the GPU suite, a C3 timing.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
{ python tools/fuzz_one.py 3871; python tools/fuzz_one.py 3237; } > $O/r02af_fuzz_cases.log 2>&1; cat $O/r02af_fuzz_cases.log | grep -a "differ\|sample"
( PBRT_FUZZ_SCENES=4000 python -m pytest tests/test_gpu.py -m gpu -q -k "fuzz" 2>&1 | grep -av "^Rendering" ) > $O/r02af_fuzz.log 2>&1; grep -a "passed\|failed\|FAILED" $O/r02af_fuzz.log | tail -5
( python -m pytest tests -m gpu -x -q 2>&1 | grep -av "^Rendering" ) > $O/r02af_pytest.log 2>&1; grep -a "passed\|failed" $O/r02af_pytest.log | tail -2
timeout 300 python tools/perf_sweep.py c3 2>&1 | tail -1
