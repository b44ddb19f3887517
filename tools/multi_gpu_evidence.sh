#!/bin/bash
# N-GPU evidence on one box (run under `gpurun --gpus N`): the bench line and the BGEMM sweep,
# one rank per GPU over NCCL, launched the way the driver launches them.
N=$1; SWEEP=${2:-full}
O=gpurun_out/ev; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
head -c 400 $O/bench_${N}gpu.json; echo
if [ "$SWEEP" = "none" ]; then exit 0; fi
if [ "$SWEEP" = "quick" ]; then export LCE_SWEEP_QUICK=1; fi
$RUN bench.py --gpus $N --workload bgemm_sweep --no-cpu-baseline > $O/bench_bgemm_sweep_${N}gpu.json 2> $O/bgemm_sweep_${N}gpu.jsonl
head -c 400 $O/bench_bgemm_sweep_${N}gpu.json; echo
grep -c '"M"' $O/bgemm_sweep_${N}gpu.jsonl
