#!/bin/bash
timeout -s KILL 200 python -m pytest tests/test_conv1x1_gpu.py -q --tb=short 2>&1 | tail -12
timeout -s KILL 200 python tests/gpu_bench_res1x1.py 64 256 2>&1 | tail -12
timeout -s KILL 200 python tests/gpu_bench_res1x1.py 256 64 2>&1 | tail -12
