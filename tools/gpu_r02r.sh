#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python tools/e2e_time.py > $O/r02r_e2e.log 2>&1; cat $O/r02r_e2e.log
