#!/bin/bash
# Round-2 GPU session W: ncu report of k_path_fast (second launch of a C3 wave) for a per-source-line instruction profile.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
export PBRT_PROFILE_WAVE=16777216
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path_fast" -s 1 -c 1 \
  -o $O/r02w_fast_d1 -f python tools/profile_render.py c3 > $O/r02w_ncu_fast.log 2>&1
tail -1 $O/r02w_ncu_fast.log; ls -la $O/r02w_fast_d1.ncu-rep
