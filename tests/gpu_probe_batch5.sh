#!/bin/bash
# round-2 batch 5 (8 GPUs): gradient exchange overlapped vs after the backward, C2 and C3
run() {  # tag, port, extra env..., -- bench args
  local tag=$1 port=$2; shift 2
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 8 --steps 20 --warmup 3 --no-sample --no-cpu-baseline "$@" > gpurun_out/bench_r02_${tag}_n8.json 2> gpurun_out/bench_r02_${tag}_n8.err
  echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_r02_${tag}_n8.json').read().strip().split('\n') if l.startswith('{')][-1])
    print('${tag}', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['ms_per_step'],2), d['config']['grad_exchange'], d['clocks'])
except Exception as e:
    print('${tag} parse failed', e); print(open('gpurun_out/bench_r02_${tag}_n8.err').read()[-1500:])
PY
}
run c2_overlap 29721 --workload c2
FDX_NO_DP_OVERLAP=1 run c2_after 29722 --workload c2
run c3_overlap 29723 --workload c3 --steps 8
FDX_NO_DP_OVERLAP=1 run c3_after 29724 --workload c3 --steps 8
FDX_GRAD_BUCKETS=2 run c2_overlap_b2 29725 --workload c2
