#!/bin/bash
mkdir -p gpurun_out
LCE_B200_LIB=build/liblce_b200_prof.so LCE_B200_TC_PROF=1 PW_SKIP_FMA=1 PW_ONLY_BIG=1 PW_REPS=1 timeout 300 python tools/pw_check.py > gpurun_out/pw_prof_13.log 2>&1
cat gpurun_out/pw_prof_13.log | grep -v "^ok" | awk 'NR<=80'
