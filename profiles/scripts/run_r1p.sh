#!/bin/bash
# r1p: attention v8 (split P.V), key compaction, packed GroupNorm apply, GEMM+col2im alpha head, dual epilogue for short-K MT=2
mkdir -p gpurun_out
bash tests/run_kernel_groups.sh 2>&1 | grep -E "===|passed|failed|error|Error|timeout"
echo "--- attention, v7c whole-tile P.V"; SDM_ATTN_SPLIT=0 timeout 120 python tests/bench_kernels.py attn 2>&1 | tee gpurun_out/kbench_attn_split0.txt
echo "--- attention, v8 split P.V";       SDM_ATTN_SPLIT=1 timeout 120 python tests/bench_kernels.py attn 2>&1 | tee gpurun_out/kbench_attn_split1.txt
for V in 0 1 2; do echo "--- SDM_GN_APPLY=$V"; SDM_GN_APPLY=$V timeout 120 python tests/bench_kernels.py "gn+" 2>&1 | tee gpurun_out/kbench_gn_v$V.txt; done
for V in 0 1; do echo "--- SDM_EWG_MT2=$V"; SDM_EWG_MT2=$V timeout 120 python tests/bench_kernels.py "conv1x1" 2>&1 | tee gpurun_out/kbench_c1_mt2_$V.txt; done
timeout 120 python tests/bench_kernels.py "conv3x3" > gpurun_out/kbench_conv_r1p.txt 2>&1; cat gpurun_out/kbench_conv_r1p.txt
runbench() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1p_$tag.csv > gpurun_out/bench_r1p_$tag.json 2> gpurun_out/bench_r1p_$tag.err
  echo "bench $tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1p_$tag.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"; tail -3 gpurun_out/bench_r1p_$tag.err
}
runbench new SDM_DUMMY=1
runbench base SDM_ATTN_SPLIT=0 SDM_ATTN_COMPACT=0 SDM_GN_APPLY=0 SDM_ALPHA_HEAD=0 SDM_EWG_MT2=0
runbench nocompact SDM_ATTN_COMPACT=0
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/engine_tests_r1p.log 2>&1
echo "engine tests exit $?"; grep -E "^\[parity|^\[compact|passed|failed|Error|error" gpurun_out/engine_tests_r1p.log | tail -20
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn8_r1p python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_attn8.log 2>&1; echo "ncu attn exit $?"
timeout 300 $NCU -k regex:gn_apply -s 4 -c 1 -o gpurun_out/prof_gnapply_r1p python tests/bench_kernels.py "gn+silu 128ch" > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn exit $?"
