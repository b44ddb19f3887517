#!/bin/bash
mkdir -p gpurun_out
PW_SKIP_FMA=1 timeout 300 python tools/pw_check.py > gpurun_out/pw_check_9.log 2>&1
tail -14 gpurun_out/pw_check_9.log
