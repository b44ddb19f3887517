#!/bin/bash
mkdir -p gpurun_out
PW_SKIP_FMA=1 PW_ONLY_BIG=1 PW_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tf32 -o gpurun_out/r02_pw_tf32 -f python tools/pw_check.py > gpurun_out/pw_ncu_10.log 2>&1
tail -5 gpurun_out/pw_ncu_10.log
ls -la gpurun_out/*.ncu-rep
