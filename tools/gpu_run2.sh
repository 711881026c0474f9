#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.log; tail -15 gpurun_out/bench.err
