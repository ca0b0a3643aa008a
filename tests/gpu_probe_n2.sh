#!/bin/bash
# two-GPU check: NCCL exchange test + the default bench line under torchrun (N=2)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_dp_nccl_gpu.py -q --tb=short 2>&1 | tail -4
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 > gpurun_out/bench_r02_all_n2.json 2> gpurun_out/bench_n2.err
tail -c 600 gpurun_out/bench_r02_all_n2.json; echo; tail -3 gpurun_out/bench_n2.err
