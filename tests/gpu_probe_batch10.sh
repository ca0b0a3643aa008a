#!/bin/bash
# round-2 batch 10: the tct tests on the rebuilt library (empty-half skip), a per-call table of one eager C2 / C3
# step (op, shape, kernel, time), and compute-sanitizer memcheck over smoke() + one C2 training step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_optin_kernels_gpu.py -x -q 2>&1 | tail -3
FDX_BENCH_CALLS=gpurun_out/calls_r02_c2.txt timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_c2_b10.json 2> gpurun_out/bench_c2_b10.err
tail -c 300 gpurun_out/bench_c2_b10.json; echo
FDX_BENCH_CALLS=gpurun_out/calls_r02_c3.txt timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 --no-sample --no-cpu-baseline > gpurun_out/bench_c3_b10.json 2> gpurun_out/bench_c3_b10.err
tail -c 300 gpurun_out/bench_c3_b10.json; echo
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_sanitizer_target.py > gpurun_out/sanitizer_r02_final.log 2>&1
echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer_r02_final.log
