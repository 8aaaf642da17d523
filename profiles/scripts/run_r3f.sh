#!/bin/bash
# r3f: Upsample2D as four polyphase 2x2-tap convs (SDM_CONV_POLY = 0 / 1), non-GroupNorm CTA-pair form; tests + step A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_parity_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
cp gpurun_out/parity_r2.json gpurun_out/r3f_parity.json 2>/dev/null
for P in 0 1 0 1; do
  SDM_CONV_POLY=$P timeout 600 python bench.py --quick --steps 4 --warmup 3 --dump-ops gpurun_out/r3f_ops_$P.csv > gpurun_out/r3f_bench_$P.json 2> gpurun_out/r3f_bench_$P.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3f_bench_$P.json'))
print('POLY=$P', 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: (v['ms'], v.get('tflops')) for k, v in list(d['kernel_breakdown'].items())[:4]}, d['kernel_breakdown'].get('tc:conv3x3_poly'))
PY
done
