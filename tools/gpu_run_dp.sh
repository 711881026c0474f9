#!/bin/bash
# data-parallel bench on N GPUs of one box (N = $1)
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 16 --warmup 12 > gpurun_out/bench_dp$N.log 2> gpurun_out/bench_dp$N.err
echo "exit $?"; tail -1 gpurun_out/bench_dp$N.log | cut -c1-700; tail -5 gpurun_out/bench_dp$N.err
