#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_sanitizer_target.py > gpurun_out/sanitizer_r02_final.log 2>&1
echo "sanitizer rc=$?"; tail -9 gpurun_out/sanitizer_r02_final.log
