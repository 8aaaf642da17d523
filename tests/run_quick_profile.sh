#!/bin/bash
# quick per-op dump + targeted ncu captures. Usage: bash tests/run_quick_profile.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-x}; KRE=${2:-gn_apply}; SKIP=${3:-3}; CNT=${4:-2}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['value'], d['ms_per_step'], d['clocks']); 
for k,v in d['kernel_breakdown'].items(): print(k, v)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o gpurun_out/prof_${TAG} -f \
    python bench.py --quick --steps 1 --warmup 0 --batch 2 > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"
