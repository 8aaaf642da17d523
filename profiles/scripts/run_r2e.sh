#!/bin/bash
# r2e: fused GroupNorm v3 (MUFU.RCP transform): micro-benchmark, kernel parity, step A/B over the fusion policy
mkdir -p gpurun_out
python tests/bench_kernels.py gn+conv 2>&1 | tee gpurun_out/r2e_kbench_gnconv.txt
( timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "fused_groupnorm or can_fuse or groupnorm" -p no:cacheprovider ) > gpurun_out/r2e_kernels.log 2>&1; echo "kernel pytest exit $?"; tail -4 gpurun_out/r2e_kernels.log
for f in 0 1 3 2; do
  SDM_GN_FUSE=$f timeout 600 python bench.py --quick --steps 4 --warmup 2 --dump-ops gpurun_out/r2e_ops_f$f.csv > gpurun_out/r2e_bench_f$f.json 2> gpurun_out/r2e_bench_f$f.err; echo "bench GN_FUSE=$f exit $?"
  python - "$f" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/r2e_bench_f{sys.argv[1]}.json'))
print('  ms/step', round(d['ms_per_step'], 2), 'clk', d['clocks']['sm_mhz'], 'conv3x3', d['kernel_breakdown']['tc:conv3x3'], 'gn', d['kernel_breakdown'].get('groupnorm'), d['kernel_breakdown'].get('groupnorm_stats'))
PY
done
