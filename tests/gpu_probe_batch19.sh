#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python tests/gpu_ncu_upconv.py
FDX_PAIR=1 timeout -s KILL 120 python tests/gpu_ncu_upconv.py
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdx_tc_kernel -s 8 -c 2 -o gpurun_out/ncu_upconv_r02 -f python tests/gpu_ncu_upconv.py > gpurun_out/ncu_upconv.log 2>&1
tail -3 gpurun_out/ncu_upconv.log
