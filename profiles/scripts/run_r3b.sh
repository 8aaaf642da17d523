#!/bin/bash
# r3b: (1) attention with one S chunk always in flight across key-tile boundaries (SDM_ATTN_PIPE=1) vs the r2z form (=0), same box;
# (2) ncu --set full of the three epilogue-heavy GEMM instantiations: fp32 scores (VAE QK^T), GEGLU, transposed V^T.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "attention or attn" 2>&1 | tail -3
for P in 0 1 0 1; do
  echo "== SDM_ATTN_PIPE=$P"
  SDM_ATTN_PIPE=$P timeout 300 python tests/bench_kernels.py attn 2>&1 | grep -E "attn_"
done
for P in 0 1 0 1; do
  SDM_ATTN_PIPE=$P timeout 600 python bench.py --quick --steps 4 --warmup 3 --dump-ops gpurun_out/r3b_ops_$P.csv > gpurun_out/r3b_bench_$P.json 2> gpurun_out/r3b_bench_$P.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3b_bench_$P.json'))
print('PIPE=$P', 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: v['ms'] for k, v in list(d['kernel_breakdown'].items())[:6]})
PY
done
for M in 3 2 1; do
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"conv_gemm_kernel<256, 1, $M," -c 2 \
      -o gpurun_out/r3b_mode$M -f python profiles/scripts/one_forward.py > gpurun_out/r3b_ncu_mode$M.log 2>&1
  tail -2 gpurun_out/r3b_ncu_mode$M.log
done
ls -la gpurun_out/*.ncu-rep
