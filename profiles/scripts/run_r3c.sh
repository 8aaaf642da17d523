#!/bin/bash
# r3c: swapped halo conv in clusters of CS CTAs with multicast weight tiles (SDM_SWH_MC = 1 / 2 / 4), after the warp-uniform issuers;
# ncu --set full of the three epilogue-heavy GEMM instantiations (fp32 scores, GEGLU, transposed V^T)
mkdir -p gpurun_out
for CS in 2 4; do
  echo "== tests SDM_SWH_MC=$CS"
  SDM_SWH_MC=$CS timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "conv" 2>&1 | tail -3
done
for CS in 1 2 4 1 2; do
  echo "== kbench SDM_SWH_MC=$CS"
  SDM_SWH_MC=$CS timeout 300 python tests/bench_kernels.py "gn+conv3x3" 2>&1 | grep -E "fused|apply"
done
echo "== prof"
SDM_GEMM_PROF=1 SDM_SWH_MC=1 timeout 120 python tests/bench_kernels.py "gn+conv3x3 256->256 @512^2 B4 fused" 2>&1 | grep -E "fused|prof" | head -8
SDM_GEMM_PROF=1 SDM_SWH_MC=2 timeout 120 python tests/bench_kernels.py "gn+conv3x3 256->256 @512^2 B4 fused" 2>&1 | grep -E "fused|prof" | head -8
for CS in 1 2 4 1 2; do
  SDM_SWH_MC=$CS timeout 600 python bench.py --quick --steps 4 --warmup 3 --dump-ops gpurun_out/r3c_ops_$CS.csv > gpurun_out/r3c_bench_$CS.json 2> gpurun_out/r3c_bench_$CS.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3c_bench_$CS.json'))
print('CS=$CS', 'ms', d['ms_per_step'], d['clocks']['sm_mhz'], {k: (v['ms'], v.get('tflops')) for k, v in list(d['kernel_breakdown'].items())[:3]})
PY
done
for M in 3 2 1; do
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"conv_gemm_kernel<\(int\)256, \(int\)1, \(int\)$M," -c 2 \
      -o gpurun_out/r3c_mode$M -f python profiles/scripts/one_forward.py > gpurun_out/r3c_ncu_mode$M.log 2>&1
  tail -1 gpurun_out/r3c_ncu_mode$M.log
done
ls -la gpurun_out/*.ncu-rep
