#!/bin/bash
# r1s: where does the CTA-pair conv kernel wait?  (SDM_GEMM_PROF cycle counters of producer / MMA issuer / epilogue)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "cta_pair" --tb=short -p no:cacheprovider 2>&1 | tail -8
for V in 0 1; do
  echo "--- SDM_PAIR=$V prof"
  SDM_PAIR=$V SDM_GEMM_PROF=1 timeout 120 python tests/bench_kernels.py "conv3x3 128->128 @1024^2 B2 +res" > gpurun_out/prof_pair$V.txt 2>&1
  grep "sdm prof" gpurun_out/prof_pair$V.txt | tail -8; grep -v "sdm prof" gpurun_out/prof_pair$V.txt | tail -3
done
for V in 0 1; do echo "--- SDM_PAIR=$V"; SDM_PAIR=$V timeout 120 python tests/bench_kernels.py "conv3x3" 2>&1 | tee gpurun_out/kbench_conv_s_pair$V.txt; done
