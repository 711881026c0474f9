#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 12 --warmup 12 > gpurun_out/bench_dp8.log 2> gpurun_out/bench_dp8.err
echo "exit $?"; tail -2 gpurun_out/bench_dp8.log | cut -c1-600; tail -12 gpurun_out/bench_dp8.err
