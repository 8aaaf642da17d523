#!/bin/bash
# r1q: attention v9 (two sequenced MMA issuers) A/B vs v7c/v8, GroupNorm apply variants, engine with compaction / GEMM+col2im alpha head
mkdir -p gpurun_out
bash tests/run_kernel_groups.sh 2>&1 | grep -E "===|passed|failed|error|Error|timeout"
for V in 1 2 3; do
  SDM_ATTN_VARIANT=$V timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=line -p no:cacheprovider 2>&1 | tail -3 | sed "s/^/[attn variant $V] /"
done
for V in 0 1 2 3; do echo "--- SDM_ATTN_VARIANT=$V"; SDM_ATTN_VARIANT=$V timeout 120 python tests/bench_kernels.py attn 2>&1 | tee gpurun_out/kbench_attn_var$V.txt; done
for V in 0 2 3; do echo "--- SDM_GN_APPLY=$V"; SDM_GN_APPLY=$V timeout 120 python tests/bench_kernels.py "gn+" 2>&1 | tee gpurun_out/kbench_gn_q$V.txt; done
runbench() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1q_$tag.csv > gpurun_out/bench_r1q_$tag.json 2> gpurun_out/bench_r1q_$tag.err
  echo "bench $tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1q_$tag.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"; tail -3 gpurun_out/bench_r1q_$tag.err
}
runbench v0 SDM_ATTN_VARIANT=0
runbench v2 SDM_ATTN_VARIANT=2
runbench v3 SDM_ATTN_VARIANT=3
runbench nocompact SDM_ATTN_VARIANT=0 SDM_ATTN_COMPACT=0
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/engine_tests_r1q.log 2>&1
echo "engine tests exit $?"; grep -E "^\[parity|^\[compact|passed|failed|Error|error" gpurun_out/engine_tests_r1q.log | tail -20
NCU="ncu --set full --clock-control none --import-source on -f"
SDM_ATTN_VARIANT=2 timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn9_r1q python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_attn9.log 2>&1; echo "ncu attn exit $?"
