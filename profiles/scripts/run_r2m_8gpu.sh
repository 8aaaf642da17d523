#!/bin/bash
# r2m (8 GPUs): BASELINE config 5 — bs=64 (8 per GPU) resolution sweep 512..1024 with the path roofline per resolution
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --sweep 512,640,768,896,1024 --steps 5 --warmup 3 > gpurun_out/r2m_sweep_8gpu.json 2> gpurun_out/r2m_sweep_8gpu.err; echo "sweep N=8 exit $?"; tail -3 gpurun_out/r2m_sweep_8gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2m_sweep_8gpu.json') if l.startswith('{')][0])
print(d['n_gpus'], d['clocks'])
for e in d['sweep']: print('  R', e['resolution'], 'global batch', e['global_batch'], 'mattes/s', round(e['value'], 2), 'ms', round(e['ms_per_step'], 2), 'frac', round(e['frac_of_sustained_peak'], 3))
PY
