#!/bin/bash
# Round-2 evidence pass on one B200 (run under gpurun). Everything lands in gpurun_out/ev/;
# tools/make_profiles.py turns it into the committed summaries under profiles/.
O=gpurun_out/ev; mkdir -p $O
NCU="ncu --clock-control none"
BENCH="python bench.py --no-cpu-baseline --no-extras"
# 1. tests + correctness tools
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 600 python tools/tc_check.py bgemm conv fused big > $O/tc_check.log 2>&1; tail -1 $O/tc_check.log
PW_SKIP_FMA=0 timeout 300 python tools/pw_check.py > $O/pw_check.log 2>&1; grep -c "^ok" $O/pw_check.log
timeout 120 python tools/stem_check.py 256 10 > $O/stem_check.log 2>&1; cat $O/stem_check.log
LCE_B200_LIB=build/liblce_b200_prof.so timeout 200 python tools/tc_prof.py > $O/tc_prof.log 2>&1
LCE_B200_LIB=build/liblce_b200_prof.so LCE_B200_TC_PROF=1 PW_SKIP_FMA=1 PW_ONLY_BIG=1 PW_REPS=1 timeout 200 python tools/pw_check.py > $O/pw_prof.log 2>&1
./tools/tc_probe > $O/tc_probe.log 2>&1
# 2. bench lines
python bench.py > $O/bench_default.json 2> $O/bench_default.err; head -c 300 $O/bench_default.json; echo
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
$BENCH --steps 1000 --warmup 10 > $O/bench_1000steps.json 2> $O/bench_1000steps.err
LCE_BENCH_VERBOSE=1 $BENCH --steps 10 --warmup 3 > $O/bench_nodes.json 2> $O/bench_nodes.err
python bench.py --workload bgemm_sweep --no-cpu-baseline > $O/bench_bgemm_sweep.json 2> $O/bgemm_sweep.jsonl
# 3. ncu: launch list of the bench command, full capture of the binary convs of one step, aux kernels, glue
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/ncu_launches.csv $BENCH --steps 2 --warmup 3 --no-e2e > $O/ncu_launches.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:bconv_tc --launch-skip 48 -c 16 -o $O/ncu_bconv_tc -f $BENCH --steps 2 --warmup 3 --no-e2e > $O/ncu_bconv_tc.log 2>&1
timeout 600 $NCU --set full -k "regex:pw_tf32|stem_conv_dw|stem7_tf32|pool2_dw3|gemm_small_m|mean_hw|softmax" --launch-skip 24 -c 12 -o $O/ncu_glue -f $BENCH --steps 2 --warmup 3 --no-e2e > $O/ncu_glue.log 2>&1
timeout 300 $NCU --set full -k "regex:bmaxpool_kernel|unpack_kernel|pack_generic_kernel|pack_f32_flat" --launch-skip 5 -c 5 -o $O/ncu_aux -f python tools/aux_kernels.py > $O/ncu_aux.log 2>&1
# gpurun copies back at most 64 MiB: export what the summaries need here, keep only the binary-conv report
for r in ncu_bconv_tc ncu_glue ncu_aux; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page details --csv > $O/${r}_details.csv 2>/dev/null
done
ncu -i $O/ncu_bconv_tc.ncu-rep --page source --csv --print-source sass --launch-skip 0 --launch-count 1 > $O/ncu_bconv_tc_s1_sass.csv 2>/dev/null
ncu -i $O/ncu_bconv_tc.ncu-rep --page source --csv --print-source sass --launch-skip 12 --launch-count 1 > $O/ncu_bconv_tc_s4_sass.csv 2>/dev/null
rm -f $O/ncu_glue.ncu-rep $O/ncu_aux.ncu-rep $O/ncu_bconv_tc.ncu-rep
timeout 300 $NCU --set full -k "regex:stem7_tf32" --launch-skip 2 -c 1 -o $O/ncu_stem7 -f python bench.py --workload birealnet18 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > $O/ncu_stem7.log 2>&1
ncu -i $O/ncu_stem7.ncu-rep --page raw --csv > $O/ncu_stem7_raw.csv 2>/dev/null; rm -f $O/ncu_stem7.ncu-rep
# 4. memcheck of the new kernels
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_builtins.py -m gpu -q -x > $O/sanitizer_builtins.log 2>&1; tail -3 $O/sanitizer_builtins.log
timeout 600 compute-sanitizer --tool memcheck python tools/tc_check.py conv fused > $O/sanitizer_tc.log 2>&1; tail -3 $O/sanitizer_tc.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/nvidia_smi.txt
ls -la $O | head -50
