#!/bin/bash
timeout -s KILL 300 python -m pytest tests/test_optin_kernels_gpu.py tests/test_tc_gpu.py -q --tb=short -x 2>&1 | tail -12
for m in fwd dgrad; do
  for v in "FDX_TCT_HALO=0" "FDX_TCT_HALO=-1"; do
    echo "== layers 64 $m $v"; env $v timeout -s KILL 200 python tests/gpu_bench_layers.py 64 256 $m 2>&1 | grep -v "^GN" | tail -23
  done
done
for v in "FDX_TCT_HALO=0" "FDX_TCT_HALO=-1" "FDX_TCT_HALO=64"; do
  echo "== bench c2 $v"
  env $v timeout -s KILL 200 python bench.py --workload c2 --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'euler', round(d['sample']['denoise_steps_per_sec'],1), d['clocks']['sm_mhz'])"
done
