#!/bin/bash
# r3a: A/B of the softmax exponentials split between MUFU.EX2 and an FMA-pipe degree-4 polynomial (SDM_ATTN_POLY = 0 none,
# 4 every fourth pair, 3 every third, 2 every second) — kernel microbench + whole step, same box.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "attention or attn" 2>&1 | tail -3
for P in 0 4 3 2 0 4; do
  echo "== SDM_ATTN_POLY=$P"
  SDM_ATTN_POLY=$P timeout 300 python tests/bench_kernels.py attn 2>&1 | grep -E "attn_" 
done
for P in 0 4 3 2 0 4; do
  SDM_ATTN_POLY=$P timeout 600 python bench.py --quick --steps 4 --warmup 3 --dump-ops gpurun_out/r3a_ops_$P.csv > gpurun_out/r3a_bench_$P.json 2> gpurun_out/r3a_bench_$P.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3a_bench_$P.json'))
print('POLY=$P', 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: v for k, v in list(d['kernel_breakdown'].items())[:6]})
PY
done
