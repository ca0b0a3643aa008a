#!/bin/bash
# round-2 batch 15: attention kernels with hoisted masks / folded scale: parity + layer numbers + C3 / C4 lines
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_attention_gpu.py tests/test_unet_gpu.py tests/test_baseline_shapes_gpu.py -q --tb=short -x 2>&1 | tail -5
echo "== attention"; timeout -s KILL 200 python tests/gpu_bench_attention.py 2>&1 | cut -c1-150
echo "== bench c3"
timeout -s KILL 300 python bench.py --workload c3 --no-cpu-baseline --no-sample --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['launches_per_step']); print({k:(round(v['ms'],2),round(v['tflops'])) for k,v in d['roofline']['kernels'].items()})"
echo "== bench c4"
timeout -s KILL 300 python bench.py --workload c4 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-900
