#!/bin/bash
python tools/profile_kind.py C4 32768 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:scl_lut_warp -s 1 -c 1 -o gpurun_out/prof_r2_c4b -f python tools/profile_kind.py C4 32768 > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/ncu_c4.log
