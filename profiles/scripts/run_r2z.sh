#!/bin/bash
# r2z (and r2i before the warp-uniform MMA issue loops): the driver's GPU test command on the round-2 default configuration
# + smoke + the default bench line
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -8 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --dump-ops gpurun_out/r2z_ops.csv > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r2z_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2z_bench.json'))
print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'gemm', d['gemm_roofline']['achieved'], 'path', d['path_roofline'])
print('worst', d['worst_case']); print('gpu_baseline', d['gpu_baseline']); print('cpu', d['cpu_baseline']); print('graph', d['config']['cuda_graph'])
for k, v in list(d['kernel_breakdown'].items())[:8]: print(' ', k, v)
PY
