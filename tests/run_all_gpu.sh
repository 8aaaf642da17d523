#!/bin/bash
# full GPU check: kernel groups, kernel micro-benchmarks, engine tests, bench. Usage: bash tests/run_all_gpu.sh <tag>
TAG=${1:-x}
bash tests/run_kernel_groups.sh
python tests/bench_kernels.py > gpurun_out/kbench_$TAG.txt 2>&1; cat gpurun_out/kbench_$TAG.txt
timeout 1500 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/engine_tests.log 2>&1
echo "engine tests exit $?"; grep -E "^\[parity|passed|failed" gpurun_out/engine_tests.log | tail -8
python bench.py --steps 4 --warmup 3 --dump-ops gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"; tail -3 gpurun_out/bench_$TAG.err
