#!/bin/bash
mkdir -p gpurun_out
python tools/stem_check.py 256 10 > gpurun_out/stem_17.log 2>&1
cat gpurun_out/stem_17.log
timeout 600 python -m pytest tests/test_gpu_builtins.py -m gpu -q -x 2>&1 | tail -3
