#!/bin/bash
# r3x: BASELINE config 2 (bs=1, 1024^2) with the e2e arm warmed, CUDA graph on / off
mkdir -p gpurun_out
for G in 1 0; do
  timeout 80 python bench.py --batch 1 --steps 10 --warmup 3 --graph $G --no-gpu-baseline --no-cpu-baseline > gpurun_out/r3x_bench_bs1_graph$G.json 2> gpurun_out/r3x_bs1_$G.err; echo "exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/r3x_bench_bs1_graph$G.json')); print('graph $G', d['value'], d['ms_per_step'], d['path_roofline']['frac_of_sustained_peak'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
