#!/bin/bash
# bench + ncu evidence for one round. Usage: bash tests/run_bench_profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
# launch list (every launch of one step; cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --quick --steps 1 --warmup 0 --batch 2 > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches_$TAG.csv
# full-set captures of the top kernels (1 GPU, few launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 4 -c 4 -o gpurun_out/prof_convgemm_$TAG -f \
    python bench.py --quick --steps 1 --warmup 0 --batch 2 > gpurun_out/ncu_conv_$TAG.log 2>&1
echo "ncu conv exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 2 -o gpurun_out/prof_attn_$TAG -f \
    python bench.py --quick --steps 1 --warmup 0 --batch 2 > gpurun_out/ncu_attn_$TAG.log 2>&1
echo "ncu attn exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_ -s 6 -c 3 -o gpurun_out/prof_gn_$TAG -f \
    python bench.py --quick --steps 1 --warmup 0 --batch 2 > gpurun_out/ncu_gn_$TAG.log 2>&1
echo "ncu gn exit $?"
ls -la gpurun_out/
