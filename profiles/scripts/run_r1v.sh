#!/bin/bash
# r1v: attention (sequenced issuers) with the FMA-pipe exp2 share re-measured; simplified GroupNorm apply / attention sources
mkdir -p gpurun_out
bash tests/run_kernel_groups.sh 2>&1 | grep -E "===|passed|failed|error|Error|timeout"
for P in 4 8; do SDM_ATTN_POLY=$P timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=line -p no:cacheprovider 2>&1 | tail -2 | sed "s/^/[poly $P] /"; done
SDM_ATTN_SPLIT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --tb=line -p no:cacheprovider 2>&1 | tail -2 | sed "s/^/[split] /"
for P in 0 4 8; do echo "--- SDM_ATTN_POLY=$P"; SDM_ATTN_POLY=$P timeout 120 python tests/bench_kernels.py attn 2>&1 | tee gpurun_out/kbench_attn_v9_poly$P.txt; done
echo "--- gn"; timeout 120 python tests/bench_kernels.py "gn+" 2>&1 | tee gpurun_out/kbench_gn_r1v.txt
for P in 0 4; do
SDM_ATTN_POLY=$P timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r1v_p$P.csv > gpurun_out/bench_r1v_p$P.json 2> gpurun_out/bench_r1v_p$P.err
echo "bench poly $P exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1v_p$P.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'])
for k,v in list(d['kernel_breakdown'].items())[:6]: print(' ', k, v)"; tail -3 gpurun_out/bench_r1v_p$P.err
done
