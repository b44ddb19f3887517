#!/bin/bash
mkdir -p gpurun_out
echo "== 2 groups"; timeout 300 python tools/tc_check.py big 2>&1 | tail -12 | cut -c1-140
echo "== 3 groups"; LCE_B200_LIB=build/liblce_b200_exp3.so timeout 300 python tools/tc_check.py bgemm conv fused big > gpurun_out/tc_check_exp3.log 2>&1; tail -12 gpurun_out/tc_check_exp3.log | cut -c1-140; grep -c "^ok" gpurun_out/tc_check_exp3.log
