#!/bin/bash
# r3y: BASELINE config 2 (bs=1, 1024^2) and the 1-GPU resolution sweep with the round-3 kernels
mkdir -p gpurun_out
timeout 100 python bench.py --batch 1 --quick --steps 10 --warmup 3 > gpurun_out/r3y_bench_bs1.json 2> gpurun_out/r3y_bench_bs1.err; echo "bs1 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r3y_bench_bs1.json')); print('bs1', d['value'], d['ms_per_step'], d['path_roofline']['frac_of_sustained_peak'], d['e2e']['ms_per_step'] if d.get('e2e') else None)"
timeout 150 python bench.py --sweep 512,640,768,896,1024 --steps 3 --warmup 3 > gpurun_out/r3y_sweep_1gpu.json 2> gpurun_out/r3y_sweep.err; echo "sweep exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r3y_sweep_1gpu.json')); print([(s['resolution'], round(s['value'],1), round(s['ms_per_step'],1), round(s['frac_of_sustained_peak'],3)) for s in d['sweep']])"
