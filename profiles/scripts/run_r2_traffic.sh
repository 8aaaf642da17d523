#!/bin/bash
# DRAM traffic of the dominant kernel family (the 3x3 convolution kernels) over one bs=8 1024^2 forward: ncu dram__bytes_read/write per
# launch -> profiles/r2_traffic.json (bench.py puts it into roofline.traffic next to the algorithmic bytes of the same launches)
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_swap|conv_gemm_kernel' \
    --csv --log-file gpurun_out/r2_traffic_launches.csv python profiles/scripts/one_forward.py > gpurun_out/r2_traffic.log 2>&1
tail -3 gpurun_out/r2_traffic.log
python profiles/scripts/reduce_traffic.py gpurun_out/r2_traffic_launches.csv gpurun_out/r2_traffic.json
