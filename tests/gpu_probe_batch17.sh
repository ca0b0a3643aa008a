#!/bin/bash
# round-2 batch 17: pixels-as-M epilogue column statistics without the 128-thread barrier: parity + per-call table
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_optin_kernels_gpu.py tests/test_tc_gpu.py tests/test_unet_gpu.py -q --tb=short -x 2>&1 | tail -6
echo "== bench c2"
FDX_BENCH_CALLS=gpurun_out/calls_c2_b17.txt timeout -s KILL 200 python bench.py --workload c2 --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'euler', round(d['sample']['denoise_steps_per_sec'],1), d['clocks']['sm_mhz'], d['launches_per_step'])"
grep upconv gpurun_out/calls_c2_b17.txt | cut -c1-140
echo "== bench c3"
FDX_BENCH_CALLS=gpurun_out/calls_c3_b17.txt timeout -s KILL 300 python bench.py --workload c3 --no-cpu-baseline --no-sample --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['launches_per_step'])"
grep upconv gpurun_out/calls_c3_b17.txt | cut -c1-140
