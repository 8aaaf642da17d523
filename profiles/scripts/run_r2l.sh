#!/bin/bash
# r2l: n3 prompts parity; BASELINE config 2 (bs=1 1024^2 latency, CUDA graph on / off); DRAM traffic of the conv kernels (ncu);
# resolution sweep on one GPU
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_prompts_gpu.py -q -s -p no:cacheprovider ) > gpurun_out/r2l_prompts.log 2>&1; echo "prompts pytest exit $?"; grep -E "^\[prompt|passed|failed|Error" gpurun_out/r2l_prompts.log | tail -12
for g in 1 0; do
  timeout 600 python bench.py --batch 1 --graph $g --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2l_bench_bs1_graph$g.json 2> gpurun_out/r2l_bench_bs1_graph$g.err; echo "bench bs=1 graph=$g exit $?"
  python - "$g" <<'PY'
import json, sys
d = json.loads([l for l in open(f'gpurun_out/r2l_bench_bs1_graph{sys.argv[1]}.json') if l.startswith('{')][0])
print('  bs=1 graph', sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), 'mattes/s', round(d['value'], 2), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'path frac', round(d['path_roofline']['frac_of_sustained_peak'], 3), 'worst', d['worst_case'] and round(d['worst_case']['ms_per_step'], 3), d['clocks'])
PY
done
timeout 900 bash profiles/scripts/run_r2_traffic.sh 2>&1 | tail -12
timeout 900 python bench.py --sweep 512,640,768,896,1024 --steps 5 --warmup 3 > gpurun_out/r2l_sweep_1gpu.json 2> gpurun_out/r2l_sweep_1gpu.err; echo "sweep exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2l_sweep_1gpu.json') if l.startswith('{')][0])
for e in d['sweep']: print('  R', e['resolution'], 'mattes/s', round(e['value'], 2), 'ms', round(e['ms_per_step'], 2), 'frac', round(e['frac_of_sustained_peak'], 3), 'launches', e['launches_per_step'])
PY
