#!/bin/bash
# round-2 batch 16: GroupNorm-backward first pass in the transposed data-gradient epilogue (default where eligible)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_optin_kernels_gpu.py tests/test_tc_gpu.py -q --tb=short -x 2>&1 | tail -8
timeout -s KILL 600 python -m pytest tests/test_unet_gpu.py tests/test_baseline_shapes_gpu.py tests/test_train_gpu.py -q --tb=short -x 2>&1 | tail -5
for v in "FDX_X=0" "FDX_GN_FUSE=0" "FDX_X=1" "FDX_GN_FUSE=0"; do
  echo "== bench c2 $v"
  env $v timeout -s KILL 200 python bench.py --workload c2 --no-cpu-baseline --no-sample --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['launches_per_step'])"
done
for v in "FDX_X=0" "FDX_GN_FUSE=0"; do
echo "== bench c3 $v"
env $v timeout -s KILL 300 python bench.py --workload c3 --no-cpu-baseline --no-sample --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['launches_per_step'])"
done
