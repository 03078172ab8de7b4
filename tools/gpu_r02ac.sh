#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
{ python tools/fuzz_one.py 3871; PBRT_B200_LIB=$PWD/pbrt-v1_b200/csrc/variants/oldwalk.so python tools/fuzz_one.py 3871; } > $O/r02ac_fuzz_3871.log 2>&1
cat $O/r02ac_fuzz_3871.log
