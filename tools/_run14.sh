#!/bin/bash
mkdir -p gpurun_out
PW_SKIP_FMA=1 timeout 300 python tools/pw_check.py > gpurun_out/pw_check_14.log 2>&1
tail -6 gpurun_out/pw_check_14.log
LCE_B200_LIB=build/liblce_b200_prof.so LCE_B200_TC_PROF=1 PW_SKIP_FMA=1 PW_ONLY_BIG=1 PW_REPS=1 timeout 300 python tools/pw_check.py > gpurun_out/pw_prof_14.log 2>&1
grep -v "^ok" gpurun_out/pw_prof_14.log | awk 'NR%16>=2 && NR%16<=9' | head -40
