#!/bin/bash
# r2c: fused GroupNorm, second version (8 branch-free transform warps, 3 halo slots / 5 weight stages) + the pruned kernel set
# (no pair / light / prefetch / split-PV variants): full kernel parity suite, then step A/B on one box
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider ) > gpurun_out/r2c_kernels.log 2>&1; echo "kernel pytest exit $?"; tail -5 gpurun_out/r2c_kernels.log
for cfg in "0 0" "1 0" "2 0"; do
  set -- $cfg
  SDM_GN_FUSE=$1 SDM_SWAP_HALO=$2 timeout 600 python bench.py --quick --steps 4 --warmup 2 --dump-ops gpurun_out/r2c_ops_f$1_h$2.csv > gpurun_out/r2c_bench_f$1_h$2.json 2> gpurun_out/r2c_bench_f$1_h$2.err; echo "bench GN_FUSE=$1 SWAP_HALO=$2 exit $?"
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/r2c_bench_f{sys.argv[1]}_h{sys.argv[2]}.json'))
print('  ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'clk', d['clocks']['sm_mhz'], 'conv3x3', d['kernel_breakdown']['tc:conv3x3'], 'gn', d['kernel_breakdown'].get('groupnorm'), d['kernel_breakdown'].get('groupnorm_stats'))
PY
done
