#!/bin/bash
# r3e: GEGLU epilogue with rcp.approx in erf (no per-element branch region), transposed-store epilogue with pipelined tcgen05.ld /
# hoisted bias; kernel tests + microbench + step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tests/bench_kernels.py "geglu" 2>&1 | grep -E "geglu"
timeout 300 python tests/bench_kernels.py "linear_T" 2>&1 | grep -E "linear_T"
for P in 1 1; do
  timeout 600 python bench.py --quick --steps 4 --warmup 3 --dump-ops gpurun_out/r3e_ops_$P.csv > gpurun_out/r3e_bench_$P.json 2> gpurun_out/r3e_bench_$P.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3e_bench_$P.json'))
print('ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: (v['ms'], v.get('tflops')) for k, v in list(d['kernel_breakdown'].items())[:9]})
PY
done
