#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_15.log 2>&1
tail -8 gpurun_out/pytest_gpu_15.log
LCE_BENCH_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_15.json 2> gpurun_out/bench_15.err
grep "^node" gpurun_out/bench_15.err | sort -k4 -n -r | head -14
python -c "
import json;d=json.load(open('gpurun_out/bench_15.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'))"
tail -3 gpurun_out/bench_15.err
