#!/bin/bash
# r1u: swapped-operand 3x3 conv kernel (channels on M, 256 pixels on N) for the 128-channel layers; in-step A/B of halo / swap
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "swapped" --tb=short -p no:cacheprovider 2>&1 | tail -25
for V in 0 1; do echo "--- SDM_SWAP=$V (SDM_HALO=0)"; SDM_HALO=0 SDM_SWAP=$V timeout 120 python tests/bench_kernels.py "conv3x3 128" 2>&1 | tee gpurun_out/kbench_conv_swap$V.txt; done
runbench() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1u_$tag.csv > gpurun_out/bench_r1u_$tag.json 2> gpurun_out/bench_r1u_$tag.err
  echo "bench $tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1u_$tag.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in list(d['kernel_breakdown'].items())[:4]: print(' ', k, v)"; tail -3 gpurun_out/bench_r1u_$tag.err
}
runbench swap1_halo0 SDM_SWAP=1 SDM_HALO=0
runbench swap0_halo0 SDM_SWAP=0 SDM_HALO=0
runbench swap1_halo1 SDM_SWAP=1 SDM_HALO=1
runbench swap0_halo1 SDM_SWAP=0 SDM_HALO=1
