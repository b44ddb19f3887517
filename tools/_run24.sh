#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_24.json 2> gpurun_out/bench_24.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_24.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'parity',d.get('parity_checked'))
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','share_of_step','avg_launch_ms')})
print('cpu',d.get('cpu_baseline'))
print('by_op',d['config'].get('by_op_ms_per_step'))
for k in ('f32_input','legacy_paths','other_configs'):
    print(k, json.dumps(d.get(k))[:600])
print('clocks',d['clocks'],'launches',d['gpu_launches'])
PY
tail -3 gpurun_out/bench_24.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_24_ref.json 2> gpurun_out/bench_24_ref.err; head -c 600 gpurun_out/bench_24_ref.json
