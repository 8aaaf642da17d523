#!/bin/bash
# r2b: GroupNorm fused into the consuming conv (conv_swap_halo_kernel<true>): kernel parity (bit-exact vs apply pass + same conv), engine
# parity with the fusion on, then step-time A/B on ONE box: no fusion / swap vs swap_halo / fusion for N=128 / fusion for all N%128==0
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "fused_groupnorm or can_fuse or swapped" -p no:cacheprovider ) > gpurun_out/r2b_kernels.log 2>&1; echo "kernel pytest exit $?"; tail -15 gpurun_out/r2b_kernels.log
( timeout 900 python -m pytest tests/test_engine_gpu.py -q -x -p no:cacheprovider ) > gpurun_out/r2b_engine.log 2>&1; echo "engine pytest exit $?"; tail -6 gpurun_out/r2b_engine.log
( SDM_GN_FUSE=2 timeout 900 python -m pytest tests/test_engine_gpu.py -q -x -k "alpha_matches or batch_independence or full_size" -p no:cacheprovider ) > gpurun_out/r2b_engine_fuse2.log 2>&1; echo "engine(fuse2) pytest exit $?"; tail -6 gpurun_out/r2b_engine_fuse2.log
for cfg in "0 0" "0 1" "1 0" "2 0"; do
  set -- $cfg
  SDM_GN_FUSE=$1 SDM_SWAP_HALO=$2 timeout 600 python bench.py --quick --steps 4 --warmup 2 --dump-ops gpurun_out/r2b_ops_f$1_h$2.csv > gpurun_out/r2b_bench_f$1_h$2.json 2> gpurun_out/r2b_bench_f$1_h$2.err; echo "bench GN_FUSE=$1 SWAP_HALO=$2 exit $?"
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/r2b_bench_f{sys.argv[1]}_h{sys.argv[2]}.json'))
print('  ms/step', round(d['ms_per_step'], 2), 'clk', d['clocks']['sm_mhz'], 'conv3x3', d['kernel_breakdown']['tc:conv3x3'], 'gn', d['kernel_breakdown'].get('groupnorm'), d['kernel_breakdown'].get('groupnorm_stats'))
PY
done
