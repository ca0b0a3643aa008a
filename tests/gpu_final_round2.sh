#!/bin/bash
# Final round-2 validation + measurement on one box: the whole -m gpu suite, compute-sanitizer over smoke() + one
# training step of every architecture, then the measurement batch (default bench line, reference arm, ncu per-kernel
# metrics for C2 / C3, SASS inventory).
mkdir -p gpurun_out
bash tests/gpu_run_all.sh
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_sanitizer_target.py > gpurun_out/sanitizer_r02_final.log 2>&1
echo "sanitizer rc=$?"; tail -4 gpurun_out/sanitizer_r02_final.log
FDX_BENCH_CALLS=gpurun_out/calls_r02_c2_final.txt timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-sample --no-cpu-baseline > /dev/null 2>&1
FDX_BENCH_CALLS=gpurun_out/calls_r02_c3_final.txt timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 --no-sample --no-cpu-baseline > /dev/null 2>&1
bash tests/gpu_measure_all.sh c3
