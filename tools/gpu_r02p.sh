#!/bin/bash
# Round-2 GPU session P: where the e2e leg's time goes; instanced scene under the general-walk variants.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python tools/e2e_time.py > $O/r02p_e2e.log 2>&1; cat $O/r02p_e2e.log
bash tools/inst_check.sh def ga ga12 gr8 ga_r8 > $O/r02p_inst.log 2>&1; cat $O/r02p_inst.log
