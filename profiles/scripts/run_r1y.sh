#!/bin/bash
# r1y: the driver's GPU test command + bench after making the conv kernel choice batch-independent
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_gpu_r1y.log 2>&1; echo "pytest -m gpu exit $?"; tail -6 gpurun_out/pytest_gpu_r1y.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --dump-ops gpurun_out/ops_r1y.csv > gpurun_out/bench_r1y.json 2> gpurun_out/bench_r1y.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r1y.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1y.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'path', d['path_roofline'])
for k,v in list(d['kernel_breakdown'].items())[:8]: print(' ', k, v)"
