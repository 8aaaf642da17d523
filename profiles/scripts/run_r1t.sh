#!/bin/bash
# r1t: resident-halo 3x3 conv kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "halo" --tb=short -p no:cacheprovider 2>&1 | tail -25
for V in 0 1; do echo "--- SDM_HALO=$V"; SDM_HALO=$V timeout 120 python tests/bench_kernels.py "conv3x3" 2>&1 | tee gpurun_out/kbench_conv_halo$V.txt; done
echo "--- SDM_HALO=1 prof"
SDM_HALO=1 SDM_GEMM_PROF=1 timeout 120 python tests/bench_kernels.py "conv3x3 128->128 @1024^2 B2 +res" > gpurun_out/prof_halo1.txt 2>&1
grep "sdm prof" gpurun_out/prof_halo1.txt | tail -6
