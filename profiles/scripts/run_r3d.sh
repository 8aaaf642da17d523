#!/bin/bash
# r3d: swapped halo conv computed by CTA pairs (cta_group::2, SDM_SWH_PAIR = 0 / 1) for the fused-GroupNorm convs with N % 256 == 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "conv" 2>&1 | tail -5
for P in 0 1 0 1; do
  echo "== kbench SDM_SWH_PAIR=$P"
  SDM_SWH_PAIR=$P timeout 300 python tests/bench_kernels.py "gn+conv3x3" 2>&1 | grep -E "256->256|512->512"
done
echo "== prof"
SDM_GEMM_PROF=1 SDM_SWH_PAIR=1 timeout 120 python tests/bench_kernels.py "gn+conv3x3 256->256 @512^2 B4 fused" 2>&1 | grep -E "fused|prof" | head -6
for P in 0 1 0 1; do
  SDM_SWH_PAIR=$P timeout 600 python bench.py --quick --steps 4 --warmup 3 --dump-ops gpurun_out/r3d_ops_$P.csv > gpurun_out/r3d_bench_$P.json 2> gpurun_out/r3d_bench_$P.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3d_bench_$P.json'))
print('PAIR=$P', 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: (v['ms'], v.get('tflops')) for k, v in list(d['kernel_breakdown'].items())[:3]})
PY
done
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
