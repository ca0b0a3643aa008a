#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"fdx_tct_kernel|fdx_wgrad9k_kernel" -s 4 -c 4 -o gpurun_out/ncu_top_r02 -f python tests/gpu_ncu_top_kernels.py > gpurun_out/ncu_top.log 2>&1
tail -3 gpurun_out/ncu_top.log; ls -la gpurun_out/ncu_top_r02.ncu-rep
