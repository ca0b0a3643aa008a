#!/bin/bash
# CTA-pair kernels (FDX_PAIR=1) against the default, interleaved, C2 and C3
for v in "FDX_X=0" "FDX_PAIR=1" "FDX_X=1" "FDX_PAIR=1"; do
  echo "== bench c2 $v"
  env $v timeout -s KILL 200 python bench.py --workload c2 --no-cpu-baseline --no-sample --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
for v in "FDX_X=0" "FDX_PAIR=1"; do
  echo "== bench c3 $v"
  env $v timeout -s KILL 300 python bench.py --workload c3 --no-cpu-baseline --no-sample --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
