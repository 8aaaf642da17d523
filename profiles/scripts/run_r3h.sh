#!/bin/bash
# r3h: residual added in the epilogue of the swapped halo conv (SDM_SWH_RES_EPI = 1) instead of identity K slices (= 0)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "conv" 2>&1 | tail -2
timeout 120 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for P in 0 1; do
  SDM_SWH_RES_EPI=$P timeout 100 python bench.py --quick --steps 3 --warmup 3 --dump-ops gpurun_out/r3h_ops_$P.csv > gpurun_out/r3h_bench_$P.json 2> gpurun_out/r3h_bench_$P.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3h_bench_$P.json'))
print('RES_EPI=$P', 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: (v['ms'], v.get('tflops')) for k, v in list(d['kernel_breakdown'].items())[:2]})
PY
done
