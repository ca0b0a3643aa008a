#!/bin/bash
# Round measurement batch (run under gpurun): bench lines for every workload, the reference arm and one
# ncu metrics pass over an eager training step of C2 (and C3 with "c3" as argument).  Outputs: gpurun_out/.
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
timeout 300 python bench.py --workload c3 --steps 5 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err
timeout 300 python bench.py --workload c4 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
timeout 300 python bench.py --workload c5 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_reference_arm.json 2> gpurun_out/bench_ref.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
# eager, one stream (no side-stream overlap): per-kernel durations are only meaningful serialised
FDX_NO_SIDE=1 timeout 600 ncu --metrics $M --clock-control none -s 2400 -c 1300 --csv --log-file gpurun_out/metrics_c2.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-sample --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
if [ "$1" = "c3" ]; then
FDX_NO_SIDE=1 timeout 900 ncu --metrics $M --clock-control none -s 2700 -c 1500 --csv --log-file gpurun_out/metrics_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-graph --no-sample --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
fi
tail -c 300 gpurun_out/bench_c2_n1.json; echo; tail -c 200 gpurun_out/bench_c3_n1.json; echo
ls -la gpurun_out | tail -14
