#!/bin/bash
# the default bench line and the reference arm (outputs: gpurun_out/bench_r02_all_n1.json, bench_r02_reference_arm.json)
mkdir -p gpurun_out
R=${FDX_ROUND:-r02}
timeout 900 python bench.py > gpurun_out/bench_${R}_all_n1.json 2> gpurun_out/bench_${R}_all_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_ref.err
tail -c 300 gpurun_out/bench_${R}_all_n1.json; echo
