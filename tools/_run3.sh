set -x
timeout 200 python tools/tc_check.py bgemm conv fused big > gpurun_out/tc_check_3.log 2>&1; echo exit=$?; grep -v "^ok" gpurun_out/tc_check_3.log | tail -20; grep "^ok" gpurun_out/tc_check_3.log | grep -E "b=256|M=4096"
LCE_B200_LIB=build/liblce_b200_prof.so timeout 100 python tools/tc_prof.py > gpurun_out/tc_prof_3.log 2>&1; grep -A8 "56x56x64 fused\|28x28x128 fused" gpurun_out/tc_prof_3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bconv_tc -s 1 -c 1 -o gpurun_out/tc_s1_fused -f python tools/tc_one.py 56 64 fused > gpurun_out/ncu_s1.log 2>&1; tail -3 gpurun_out/ncu_s1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bconv_tc -s 1 -c 1 -o gpurun_out/tc_s3_fused -f python tools/tc_one.py 14 256 fused > gpurun_out/ncu_s3.log 2>&1; tail -3 gpurun_out/ncu_s3.log
LCE_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_3.json 2> gpurun_out/bench_3.err; echo bench_exit=$?; tail -c 3000 gpurun_out/bench_3.json; grep "^node" gpurun_out/bench_3.err | head -50; tail -5 gpurun_out/bench_3.err
