#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_builtins.py -m gpu -q -x -k "7x7" 2>&1 | tail -5
LCE_BENCH_VERBOSE=1 python bench.py --workload birealnet18 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_28.json 2> gpurun_out/bench_28.err
grep "^node" gpurun_out/bench_28.err | sort -k4 -n -r | head -8
python -c "
import json;d=json.load(open('gpurun_out/bench_28.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'], d.get('parity_checked'))"
tail -3 gpurun_out/bench_28.err
