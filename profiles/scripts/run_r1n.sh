#!/bin/bash
# r1n: new EPI_F16 epilogue + attention v7b (event-driven MMA issue): all kernel tests, micro-benchmarks, bench, ncu
mkdir -p gpurun_out
bash tests/run_kernel_groups.sh 2>&1 | grep -E "===|passed|failed|error" 
python tests/bench_kernels.py > gpurun_out/kbench_r1n.txt 2>&1; cat gpurun_out/kbench_r1n.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn7b_r1n python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 300 $NCU -k regex:conv_gemm_kernel -s 4 -c 1 -o gpurun_out/prof_c3x3_r1n python tests/bench_kernels.py "128->128 @1024^2 B2 +res" > gpurun_out/ncu_c3.log 2>&1; echo "ncu 3x3 exit $?"
timeout 300 $NCU -k regex:conv_gemm_kernel -s 4 -c 1 -o gpurun_out/prof_c1x1_r1n python tests/bench_kernels.py "im2col" > gpurun_out/ncu_c1.log 2>&1; echo "ncu 1x1 exit $?"
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s -k "alpha_matches_oracle or golden or batch_independence" > gpurun_out/engine_tests.log 2>&1
echo "engine tests exit $?"; grep -E "^\[parity|passed|failed|Error" gpurun_out/engine_tests.log | tail -12
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1n.csv > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1n.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"; tail -3 gpurun_out/bench_r1n.err
