#!/bin/bash
# r3g: residual K slices interleaved with the input slices in the swapped halo conv; kernel + engine tests, step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tests/bench_kernels.py "+res" 2>&1 | grep -E "res"
for P in 1 2; do
  timeout 600 python bench.py --quick --steps 4 --warmup 3 --dump-ops gpurun_out/r3g_ops_$P.csv > gpurun_out/r3g_bench_$P.json 2> gpurun_out/r3g_bench_$P.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3g_bench_$P.json'))
print('ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: (v['ms'], v.get('tflops')) for k, v in list(d['kernel_breakdown'].items())[:4]})
PY
done
