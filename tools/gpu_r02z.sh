#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python tools/host_gap.py 8 > $O/r02z_host_gap.log 2>&1; PBRT_HOST_PHASES=1 python tools/host_gap.py 8 2>&1 | tail -60 >> $O/r02z_host_gap.log; cat $O/r02z_host_gap.log
