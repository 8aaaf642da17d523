#!/bin/bash
# r1o: attention v7c (split S halves, S-priority issue, optional FMA-pipe exp2), GEGLU/F32 pipelined epilogues, two-stage GN reduce
mkdir -p gpurun_out
bash tests/run_kernel_groups.sh 2>&1 | grep -E "===|passed|failed|error|Error" 
for P in 0 4 6 8; do echo "--- SDM_ATTN_POLY=$P"; SDM_ATTN_POLY=$P python tests/bench_kernels.py attn 2>&1 | tee gpurun_out/kbench_attn_poly$P.txt; done
SDM_ATTN_POLY=6 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=short -p no:cacheprovider 2>&1 | tail -3
python tests/bench_kernels.py "conv" > gpurun_out/kbench_conv_r1o.txt 2>&1; cat gpurun_out/kbench_conv_r1o.txt
python tests/bench_kernels.py "linear" > gpurun_out/kbench_lin_r1o.txt 2>&1; cat gpurun_out/kbench_lin_r1o.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn7c_r1o python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
SDM_ATTN_POLY=6 timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn7c_poly6_r1o python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_attn6.log 2>&1; echo "ncu attn poly6 exit $?"
for P in 0 6; do
SDM_ATTN_POLY=$P python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1o_p$P.csv > gpurun_out/bench_r1o_p$P.json 2> gpurun_out/bench_r1o_p$P.err
echo "bench poly=$P exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1o_p$P.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"; tail -3 gpurun_out/bench_r1o_p$P.err
done
SDM_ATTN_POLY=6 timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s -k "alpha_matches_oracle or golden or batch_independence" > gpurun_out/engine_tests_p6.log 2>&1
echo "engine tests (poly 6) exit $?"; grep -E "^\[parity|passed|failed|Error" gpurun_out/engine_tests_p6.log | tail -12
