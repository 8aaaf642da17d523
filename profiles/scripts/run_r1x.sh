#!/bin/bash
# r1x: round-end evidence: the driver's GPU test command, smoke(), bench (both arms), ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv | tee gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_gpu_r1x.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu_r1x.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --dump-ops gpurun_out/ops_r1x.csv > gpurun_out/bench_r1x.json 2> gpurun_out/bench_r1x.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r1x.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1x.json')); print('VALUE', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'path', d['path_roofline'])
for k,v in d['kernel_breakdown'].items(): print(' ', k, v)"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r1x.json 2> gpurun_out/bench_ref_r1x.err; echo "reference arm exit $?"; cat gpurun_out/bench_ref_r1x.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1x.csv \
    python bench.py --quick --steps 1 --warmup 0 --batch 2 > gpurun_out/ncu_list_r1x.log 2>&1; echo "ncu list exit $?"; wc -l gpurun_out/launches_r1x.csv
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_kernel -s 4 -c 1 -o gpurun_out/prof_attn_r1x python tests/bench_kernels.py "attn_cross_L0" > gpurun_out/ncu_a.log 2>&1; echo "ncu attn exit $?"
timeout 300 $NCU -k regex:conv_swap_kernel -s 4 -c 1 -o gpurun_out/prof_convswap_r1x python tests/bench_kernels.py "conv3x3 128->128 @1024^2 B2 +res" > gpurun_out/ncu_b.log 2>&1; echo "ncu swap exit $?"
timeout 300 $NCU -k regex:conv_gemm_kernel -s 4 -c 1 -o gpurun_out/prof_convhalo_r1x python tests/bench_kernels.py "conv3x3 512->512" > gpurun_out/ncu_c.log 2>&1; echo "ncu conv halo exit $?"
timeout 300 $NCU -k regex:gn_apply -s 4 -c 1 -o gpurun_out/prof_gnapply_r1x python tests/bench_kernels.py "gn+silu 128ch" > gpurun_out/ncu_d.log 2>&1; echo "ncu gn exit $?"
python tests/bench_kernels.py > gpurun_out/kbench_r1x.txt 2>&1; cat gpurun_out/kbench_r1x.txt
