#!/bin/bash
# round-2 batch 12: eight epilogue warps in the transposed convolution kernel: parity, layer table, bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_tc_gpu.py tests/test_optin_kernels_gpu.py tests/test_kernels_gpu.py -q --tb=short -x 2>&1 | tail -6
timeout -s KILL 600 python -m pytest tests/test_unet_gpu.py tests/test_baseline_shapes_gpu.py -q --tb=short -x 2>&1 | tail -5
for m in fwd dgrad; do
  echo "== layers 64 $m"; timeout -s KILL 200 python tests/gpu_bench_layers.py 64 256 $m 2>&1 | grep -v "^GN"
done
for v in "FDX_X=0"; do
  echo "== bench c2 $v"
  env $v FDX_BENCH_CALLS=gpurun_out/calls_c2_b12.txt timeout -s KILL 200 python bench.py --workload c2 --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'euler', round(d['sample']['denoise_steps_per_sec'],1), d['clocks']['sm_mhz'], d['launches_per_step'])"
done
echo "== bench c3"
FDX_BENCH_CALLS=gpurun_out/calls_c3_b12.txt timeout -s KILL 300 python bench.py --workload c3 --no-cpu-baseline --no-sample --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['launches_per_step']); print({k:(round(v['ms'],2),round(v['tflops'])) for k,v in d['roofline']['kernels'].items()})"
