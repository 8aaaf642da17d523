#!/bin/bash
# r1l: attention v5 vs v6 A/B, epilogue-bound GEMM cases, ncu source-level captures of the three suspects
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=short -p no:cacheprovider > gpurun_out/kern_attn.log 2>&1
echo "attention tests exit $?"; tail -5 gpurun_out/kern_attn.log
SDM_ATTN=5 python tests/bench_kernels.py attn > gpurun_out/kbench_attn5.txt 2>&1; cat gpurun_out/kbench_attn5.txt
SDM_ATTN=6 python tests/bench_kernels.py attn > gpurun_out/kbench_attn6.txt 2>&1; cat gpurun_out/kbench_attn6.txt
python tests/bench_kernels.py conv > gpurun_out/kbench_conv.txt 2>&1; cat gpurun_out/kbench_conv.txt
python tests/bench_kernels.py linear > gpurun_out/kbench_lin.txt 2>&1; cat gpurun_out/kbench_lin.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn6_r1l python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn6b_r1l python tests/bench_kernels.py "attn_self_L0" > gpurun_out/ncu_attnb.log 2>&1; echo "ncu attn-bias exit $?"
timeout 300 $NCU -k regex:conv_gemm_kernel -s 4 -c 1 -o gpurun_out/prof_c1x1_r1l python tests/bench_kernels.py "im2col" > gpurun_out/ncu_c1.log 2>&1; echo "ncu 1x1 exit $?"
timeout 300 $NCU -k regex:conv_gemm_kernel -s 4 -c 1 -o gpurun_out/prof_c3x3_r1l python tests/bench_kernels.py "128->128 @1024^2 B2 +res" > gpurun_out/ncu_c3.log 2>&1; echo "ncu 3x3 exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1l.csv > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1l.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"; tail -3 gpurun_out/bench_r1l.err
ls -la gpurun_out | head -40
