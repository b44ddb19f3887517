#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_conv_dw -c 1 -o gpurun_out/r02_stem_v5 -f python tools/stem_check.py 256 1 > gpurun_out/stem_ncu_20.log 2>&1
tail -3 gpurun_out/stem_ncu_20.log
