timeout 200 python tools/tc_check.py bgemm conv fused big > gpurun_out/tc_check_6.log 2>&1; echo exit=$?; grep -v "^ok" gpurun_out/tc_check_6.log | tail -20; grep "^ok" gpurun_out/tc_check_6.log | grep -E "b=256|M=4096"
LCE_B200_LIB=build/liblce_b200_prof.so timeout 100 python tools/tc_prof.py > gpurun_out/tc_prof_6.log 2>&1; grep -A9 "56x56x64 fused\|14x14x256 fused\|bgemm 4096x4096x8192" gpurun_out/tc_prof_6.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_6.log 2>&1; echo pytest_exit=$?; tail -25 gpurun_out/pytest_gpu_6.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_6.json 2> gpurun_out/bench_6.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_6.json').read().strip().splitlines()[-1]); print('quicknet', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'], d['config'].get('by_op_ms_per_step'))"
