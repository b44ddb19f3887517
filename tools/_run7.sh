#!/bin/bash
mkdir -p gpurun_out
./tools/tc_probe > gpurun_out/tc_probe_7.log 2>&1
tail -4 gpurun_out/tc_probe_7.log
LCE_BENCH_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_7.json 2> gpurun_out/bench_7.err
grep "^node" gpurun_out/bench_7.err | sort -k4 -n -r | head -20
