#!/bin/bash
# round-2 batch 3: measurements (default bench line, reference arm, ncu per-kernel metrics, SASS)
bash tests/gpu_measure_all.sh c3
timeout 120 python -m pytest tests/test_optin_kernels_gpu.py -q -k single_launch 2>&1 | tail -3
