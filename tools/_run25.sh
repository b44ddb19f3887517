#!/bin/bash
mkdir -p gpurun_out
for sh in "56 64" "28 128" "14 256" "7 512"; do
  set -- $sh
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:bconv_tc --launch-skip 2 -c 1 -o gpurun_out/r02_tc_s$1_fused -f python tools/tc_one.py $1 $2 fused > gpurun_out/tc_ncu_$1.log 2>&1
  tail -1 gpurun_out/tc_ncu_$1.log
done
ls -la gpurun_out/r02_tc_s*.ncu-rep
