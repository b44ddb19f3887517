#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
LCE_BENCH_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_35.json 2> gpurun_out/bench_35.err
grep "^node" gpurun_out/bench_35.err | tail -4
python -c "
import json;d=json.load(open('gpurun_out/bench_35.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'])"
