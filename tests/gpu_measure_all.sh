#!/bin/bash
# Round measurement batch (run under gpurun): the default bench line (64^2 + 256^2 blocks), the reference arm,
# an ncu metrics pass (time, DRAM bytes, tensor-pipe activity per launch) over ONE eager training step of C2
# (and of C3 with "c3" as argument), and the SASS inventory of the shipped library.  Outputs: gpurun_out/.
mkdir -p gpurun_out
R=${FDX_ROUND:-r02}
timeout 900 python bench.py > gpurun_out/bench_${R}_all_n1.json 2> gpurun_out/bench_${R}_all_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_ref.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
# eager, one stream (no side-stream overlap): per-kernel durations are only meaningful serialised.
# -s skips the warm-up steps' launches (3 eager + capture), -c bounds the capture to a little over one step.
FDX_NO_SIDE=1 timeout 900 ncu --metrics $M --clock-control none -s ${FDX_NCU_SKIP:-2200} -c ${FDX_NCU_COUNT:-1200} --csv \
    --log-file gpurun_out/metrics_${R}_train_c2_eager.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-graph --no-sample --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
if [ "$1" = "c3" ]; then
FDX_NO_SIDE=1 timeout 1200 ncu --metrics $M --clock-control none -s ${FDX_NCU_SKIP3:-2400} -c ${FDX_NCU_COUNT3:-1300} --csv \
    --log-file gpurun_out/metrics_${R}_train_c3_eager.csv \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-graph --no-sample --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
fi
cuobjdump -sass flaxdiff_b200/lib/libfdx.so 2>/dev/null | grep -oE "UTCHMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|LDTM[.A-Z0-9_x]*|STTM[.A-Z0-9_x]*|UTCBAR[.A-Z0-9_]*|FFMA2|FADD2|FMUL2|REDG[.A-Z0-9_]*|UTCCP[.A-Z0-9_]*" \
    | sort | uniq -c | sort -rn > gpurun_out/sass_${R}_summary.txt
tail -c 400 gpurun_out/bench_${R}_all_n1.json; echo
ls -la gpurun_out | tail -8
