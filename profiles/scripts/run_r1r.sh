#!/bin/bash
# r1r: CTA-pair (cta_group::2) conv kernel for the 256x128 tiles, attention v9 default, engine parity
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "cta_pair" --tb=short -p no:cacheprovider 2>&1 | tail -15
bash tests/run_kernel_groups.sh 2>&1 | grep -E "===|passed|failed|error|Error|timeout"
for V in 0 1; do echo "--- SDM_PAIR=$V"; SDM_PAIR=$V timeout 120 python tests/bench_kernels.py "conv3x3" 2>&1 | tee gpurun_out/kbench_conv_pair$V.txt; done
runbench() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1r_$tag.csv > gpurun_out/bench_r1r_$tag.json 2> gpurun_out/bench_r1r_$tag.err
  echo "bench $tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1r_$tag.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"; tail -3 gpurun_out/bench_r1r_$tag.err
}
runbench pair1 SDM_PAIR=1
runbench pair0 SDM_PAIR=0
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/engine_tests_r1r.log 2>&1
echo "engine tests exit $?"; grep -E "^\[parity|^\[compact|passed|failed|Error|error" gpurun_out/engine_tests_r1r.log | tail -20
NCU="ncu --set full --clock-control none --import-source on -f"
SDM_PAIR=1 timeout 300 $NCU -k regex:conv_gemm_kernel -s 4 -c 1 -o gpurun_out/prof_c3x3_pair_r1r python tests/bench_kernels.py "conv3x3 128->128 @1024^2 B2 +res" > gpurun_out/ncu_c3p.log 2>&1; echo "ncu conv pair exit $?"
