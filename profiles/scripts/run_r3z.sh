#!/bin/bash
# r3z: round-end validation of the r3 state: the driver's GPU test command, smoke, the default bench line, the ncu launch list
mkdir -p gpurun_out
if ! timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -p no:cacheprovider -k "batch_independence" > gpurun_out/r3z_determinism.log 2>&1; then
  echo "DETERMINISM FAILED with the interleaved residual: continuing with SDM_SWH_MIX=0"; tail -5 gpurun_out/r3z_determinism.log
  export SDM_SWH_MIX=0
else
  echo "determinism ok (SDM_SWH_MIX default)"
fi
( time timeout 480 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r3z_pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -6 gpurun_out/r3z_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 330 python bench.py --dump-ops gpurun_out/r3z_ops.csv > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r3z_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r3z_bench.json'))
print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'gemm', d['gemm_roofline']['achieved'], 'path', d['path_roofline'])
print('worst', d['worst_case']); print('gpu_baseline', d['gpu_baseline']); print('cpu', d['cpu_baseline']); print('graph', d['config']['cuda_graph'])
for k, v in list(d['kernel_breakdown'].items())[:10]: print(' ', k, v)
PY
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3z_launches.csv python bench.py --quick --steps 1 --warmup 1 --graph 0 > gpurun_out/r3z_ncu_bench.log 2>&1; echo "ncu exit $?"; wc -l gpurun_out/r3z_launches.csv
