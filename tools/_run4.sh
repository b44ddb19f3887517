timeout 60 ./tools/tc_probe 2>&1 | grep -A12 "T5" > gpurun_out/tc_probe_4.log; cat gpurun_out/tc_probe_4.log
timeout 200 python tools/tc_check.py bgemm conv fused big > gpurun_out/tc_check_4.log 2>&1; echo exit=$?; grep -v "^ok" gpurun_out/tc_check_4.log | tail -20; grep "^ok" gpurun_out/tc_check_4.log | grep -E "b=256|M=4096"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_4.json 2> gpurun_out/bench_4.err; echo bench_exit=$?; python -c "
import json; d=json.loads(open('gpurun_out/bench_4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'], d['config'].get('by_op_ms_per_step'))"
