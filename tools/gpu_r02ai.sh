#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
{ timeout 100 python tools/big_scene_check.py c4 --full; timeout 150 python tools/big_scene_check.py c5 --full; } 2>&1 | grep -av "^Rendering" > $O/r02ai_crop_parity.log; grep -a "parity\|full frame\|prims" $O/r02ai_crop_parity.log
