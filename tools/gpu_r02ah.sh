#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 100 python tools/tile_fuzz.py 1500 0 2>&1 | grep -av "^Rendering\|Warning\|Line " | tail -12 > $O/r02ah_tile_fuzz.log; cat $O/r02ah_tile_fuzz.log
