#!/bin/bash
# r1m: attention v7 (P in TMEM, TS-mode PV) parity + micro-benchmark + ncu; pre/post kernels parity; node e2e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=short -p no:cacheprovider > gpurun_out/kern_attn.log 2>&1
echo "attention tests exit $?"; tail -15 gpurun_out/kern_attn.log
timeout 600 python -m pytest tests/test_prepost_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/prepost.log 2>&1
echo "prepost tests exit $?"; grep -vE "^\[prepost\]" gpurun_out/prepost.log | tail -25; grep -E "^\[prepost\]" gpurun_out/prepost.log | sort -t= -k4 | tail -5
python tests/bench_kernels.py attn > gpurun_out/kbench_attn7.txt 2>&1; cat gpurun_out/kbench_attn7.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn7_r1m python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s -k "node_end_to_end or alpha_matches_oracle or golden" > gpurun_out/engine_tests.log 2>&1
echo "engine tests exit $?"; grep -E "^\[parity|^\[node|passed|failed|Error" gpurun_out/engine_tests.log | tail -12
