#!/bin/bash
mkdir -p gpurun_out
python tools/stem_check.py 256 10 > gpurun_out/stem_16.log 2>&1
cat gpurun_out/stem_16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_conv_dw -c 1 -o gpurun_out/r02_stem -f python tools/stem_check.py 256 1 > gpurun_out/stem_ncu_16.log 2>&1
tail -3 gpurun_out/stem_ncu_16.log
