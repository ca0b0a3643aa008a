#!/bin/bash
# round-2 batch 4 (2 GPUs): NCCL exchange test, then the default bench line under torchrun
timeout -s KILL 400 python -m pytest tests/test_dp_nccl_gpu.py -q --tb=short 2>&1 | tail -25
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_all_n2.json 2> gpurun_out/bench_r02_all_n2.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r02_all_n2.json; tail -5 gpurun_out/bench_r02_all_n2.err
FDX_NO_DP_OVERLAP=1 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 10 --warmup 3 --workload c2 --no-sample --no-cpu-baseline > gpurun_out/bench_r02_c2_n2_nooverlap.json 2> gpurun_out/bench_r02_c2_n2_nooverlap.err
echo "bench(no overlap) rc=$?"; head -c 400 gpurun_out/bench_r02_c2_n2_nooverlap.json
