#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_check.py bgemm conv fused big > gpurun_out/tc_check_26.log 2>&1; grep -c "^ok" gpurun_out/tc_check_26.log; grep "FAIL" gpurun_out/tc_check_26.log | head; tail -12 gpurun_out/tc_check_26.log | cut -c1-150
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_26.log 2>&1; tail -3 gpurun_out/pytest_gpu_26.log
LCE_BENCH_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_26.json 2> gpurun_out/bench_26.err
grep "^node" gpurun_out/bench_26.err | sort -k4 -n -r | head -8
python -c "
import json;d=json.load(open('gpurun_out/bench_26.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print(d['config']['by_op_ms_per_step'])"
