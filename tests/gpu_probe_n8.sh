#!/bin/bash
# eight-GPU check of the final library: C2 and C3 training under torchrun (no sampling blocks, no CPU baseline)
mkdir -p gpurun_out
N=${FDX_N:-8}
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c2 --no-sample --no-cpu-baseline --steps 20 > gpurun_out/bench_r02_c2_final_n$N.json 2> gpurun_out/bench_c2_n$N.err
tail -c 200 gpurun_out/bench_r02_c2_final_n$N.json; echo
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload c3 --no-sample --no-cpu-baseline --steps 5 > gpurun_out/bench_r02_c3_final_n$N.json 2> gpurun_out/bench_c3_n$N.err
tail -c 200 gpurun_out/bench_r02_c3_final_n$N.json; echo
