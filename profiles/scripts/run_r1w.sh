#!/bin/bash
# r1w: kernel groups after the clean-up, vae_qk cycle counters, halo on/off in the step (5 timed steps each, same box)
mkdir -p gpurun_out
bash tests/run_kernel_groups.sh 2>&1 | grep -E "===|passed|failed|error|Error|timeout"
SDM_GEMM_PROF=1 timeout 120 python tests/bench_kernels.py "vae_qk" > gpurun_out/prof_vaeqk.txt 2>&1
grep "sdm prof" gpurun_out/prof_vaeqk.txt | tail -8; grep -v "sdm prof" gpurun_out/prof_vaeqk.txt | tail -2
for H in 0 1 0 1; do
SDM_HALO=$H timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1w_h$H.json 2> gpurun_out/bench_r1w_h$H.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1w_h$H.json')); print('HALO=$H VALUE', round(d['value'],3), 'ms', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],3))"; tail -2 gpurun_out/bench_r1w_h$H.err
done
