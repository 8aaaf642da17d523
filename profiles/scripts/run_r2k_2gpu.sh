#!/bin/bash
# r2k (2 GPUs): the N-way sharded node call is bit-identical to the single-GPU call (one process, one handle + thread per GPU);
# then the bench at N=2 under torchrun (NCCL all-gather of alpha)
mkdir -p gpurun_out
nvidia-smi -L
( timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -p no:cacheprovider ) > gpurun_out/r2k_multigpu.log 2>&1; echo "multigpu pytest exit $?"; tail -5 gpurun_out/r2k_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo "bench N=2 exit $?"; tail -2 gpurun_out/r2k_bench_2gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2k_bench_2gpu.json') if l.startswith('{')][0])
print('N=2 VALUE', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
PY
