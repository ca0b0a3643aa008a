#!/bin/bash
# round-2 batch 13: 1x1 residual convolutions, flat GEMM launch vs conv geometry (transposed engine, 8 epilogue warps)
timeout -s KILL 120 python -m pytest tests/test_conv1x1_gpu.py -q 2>&1 | tail -2
timeout -s KILL 200 python tests/gpu_bench_res1x1.py 64 256 2>&1 | cut -c1-200
timeout -s KILL 200 python tests/gpu_bench_res1x1.py 256 64 2>&1 | cut -c1-200
