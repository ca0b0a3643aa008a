#!/bin/bash
# zero-code A/Bs on the final library: GroupNorm grid waves, fused column statistics off
for v in "FDX_X=0" "FDX_GN_WAVES=3" "FDX_GN_WAVES=4" "FDX_GN_WAVES=1" "FDX_NO_COLSTATS=1" "FDX_X=1"; do
  echo "== bench c2 $v"
  env $v timeout -s KILL 200 python bench.py --workload c2 --no-cpu-baseline --no-sample --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
