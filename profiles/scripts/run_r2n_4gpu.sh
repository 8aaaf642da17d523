#!/bin/bash
# r2n (4 GPUs): BASELINE config 4 — bs=32 768x768 batch-sharded over 4 GPUs (8 per GPU) with the NCCL all-gather of alpha
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --size 768 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2n_bench_768_4gpu.json 2> gpurun_out/r2n_bench_768_4gpu.err; echo "bench 768 N=4 exit $?"; tail -3 gpurun_out/r2n_bench_768_4gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2n_bench_768_4gpu.json') if l.startswith('{')][0])
print('N=4 768^2 bs=32: mattes/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'path', d['path_roofline'], d['clocks'])
PY
