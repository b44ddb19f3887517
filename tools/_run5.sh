timeout 60 ./tools/tc_probe 2>&1 | grep -A10 "T6" > gpurun_out/tc_probe_5.log; cat gpurun_out/tc_probe_5.log
timeout 200 python tools/tc_check.py conv fused big > gpurun_out/tc_check_5.log 2>&1; echo exit=$?; grep -v "^ok" gpurun_out/tc_check_5.log | tail -20; grep "^ok" gpurun_out/tc_check_5.log | grep -E "b=256|zp=1"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_5.log 2>&1; echo pytest_exit=$?; tail -15 gpurun_out/pytest_gpu_5.log
LCE_B200_FUSE_CONV_QUANT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_5q.json 2> gpurun_out/bench_5q.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_5q.json').read().strip().splitlines()[-1]); print('fuse_conv_quant', d['value'], d['ms_per_step'], d['config'].get('by_op_ms_per_step'))"
timeout 300 python bench.py --workload birealnet18 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_5b.json 2> gpurun_out/bench_5b.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_5b.json').read().strip().splitlines()[-1]); print('birealnet', d['value'], d['ms_per_step'], d['config'].get('by_op_ms_per_step'))"
