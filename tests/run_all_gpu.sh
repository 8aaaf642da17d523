#!/bin/bash
# full GPU check: kernel groups, engine tests, bench. Usage: bash tests/run_all_gpu.sh <tag>
TAG=${1:-x}
bash tests/run_kernel_groups.sh
timeout 1500 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/engine_tests.log 2>&1
echo "engine tests exit $?"; grep -E "^\[|passed|failed" gpurun_out/engine_tests.log | tail -25
python bench.py --steps 4 --warmup 3 --dump-ops gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
