#!/bin/bash
# round-2 batch 14: residual 1x1 through conv geometry in the UNet (parity + bench), ncu --set full of the attention kernels
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_unet_gpu.py tests/test_baseline_shapes_gpu.py tests/test_train_gpu.py -q --tb=short -x 2>&1 | tail -5
for v in "FDX_RES1X1_GEMM=1" "FDX_X=0"; do
  echo "== bench c2 $v"
  env $v timeout -s KILL 200 python bench.py --workload c2 --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'euler', round(d['sample']['denoise_steps_per_sec'],1), d['clocks']['sm_mhz'], d['launches_per_step'])"
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fdx_attn -s 4 -c 4 -o gpurun_out/ncu_attn_r02 -f python tests/gpu_ncu_attention.py > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log; ls -la gpurun_out/ncu_attn_r02.ncu-rep
