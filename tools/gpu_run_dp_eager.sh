#!/bin/bash
# eager (no CUDA graph) data-parallel step: bucketed all-reduces overlapped with backward
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --no-graphs --no-cpu > gpurun_out/bench_dp_eager.log 2> gpurun_out/bench_dp_eager.err
echo "exit $?"; tail -1 gpurun_out/bench_dp_eager.log | cut -c1-260; tail -3 gpurun_out/bench_dp_eager.err | cut -c1-200
