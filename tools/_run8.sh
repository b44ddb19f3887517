#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pw_check.py > gpurun_out/pw_check_8.log 2>&1
tail -40 gpurun_out/pw_check_8.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_8.log 2>&1
tail -5 gpurun_out/pytest_gpu_8.log
LCE_BENCH_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_8.json 2> gpurun_out/bench_8.err
grep "^node" gpurun_out/bench_8.err | sort -k4 -n -r | head -12
python -c "
import json;d=json.load(open('gpurun_out/bench_8.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'))"
