#!/bin/bash
# r2a: first GPU call of round 2 — (1) the staged swapped-operand halo kernel of r1 on hardware for the first time, (2) parity at the
# BASELINE sizes against the GPU-resident fp32 / autocast checkers + per-block error growth, (3) node call / CUDA graph / multi-device
# tests, (4) the bench line with gpu_baseline / worst_case / e2e-through-the-node; the CPU oracle's real 1024^2 rate is measured on the
# host cores in the background while the GPU tests run (it only needs the CPU).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | head -2
nproc
( python - <<'PY' > gpurun_out/r2a_cpu1024.log 2>&1
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
th = os.cpu_count() or 1
r5, t5, n5, d5 = bench.cpu_oracle_rate(512, 2, th)
r10, t10, n10, d10 = bench.cpu_oracle_rate(1024, 1, th)
json.dump({"cores": th, "seconds_per_matte_512": t5 / n5, "seconds_per_matte_1024": t10 / n10, "desc_512": d5, "desc_1024": d10,
           "note": "measured while GPU tests of the same box were running (they use one host thread)"}, open("gpurun_out/r2a_cpu_1024.json", "w"), indent=1)
PY
) &
CPU_PID=$!
( SDM_SWAP_HALO=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k swapped -p no:cacheprovider ) > gpurun_out/r2a_swap_halo.log 2>&1; echo "swap_halo pytest exit $?"; tail -5 gpurun_out/r2a_swap_halo.log
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -q -s -p no:cacheprovider ) > gpurun_out/r2a_parity.log 2>&1; echo "parity pytest exit $?"; grep -E "passed|failed|Error" gpurun_out/r2a_parity.log | tail -8
( time timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_engine_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2a_engine.log 2>&1; echo "engine pytest exit $?"; tail -6 gpurun_out/r2a_engine.log
wait $CPU_PID; cat gpurun_out/r2a_cpu_1024.json; tail -3 gpurun_out/r2a_cpu1024.log
timeout 1200 python bench.py --dump-ops gpurun_out/r2a_ops.csv > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2a_bench.json'))
print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'path', d['path_roofline'])
print('worst', d['worst_case']); print('gpu_baseline', d['gpu_baseline']); print('cpu', d['cpu_baseline']); print('graph', d['config']['cuda_graph'])
for k, v in list(d['kernel_breakdown'].items())[:12]: print(' ', k, v)
PY
