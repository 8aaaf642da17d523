#!/bin/bash
# Run the kernel parity tests group by group in separate processes (a device-side trap in one group
# must not hide the results of the others). Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv | tee gpurun_out/gpu.txt
python -c "import os; print('host cpus', os.cpu_count())" | tee -a gpurun_out/gpu.txt
i=0
for k in "linear or geglu or scores or light" "conv3x3 or conv1x1 or conv_epilogue or skinny or alpha or cta_pair or halo or swapped" "attention or key_compact" "norm or softmax or small_c"; do
  i=$((i+1))
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$k" --tb=line -p no:cacheprovider > gpurun_out/kern_$i.log 2>&1
  echo "=== group '$k' exit $?"; tail -25 gpurun_out/kern_$i.log
done
