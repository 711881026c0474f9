#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 -x -k "implicit or model or golden" > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -30 gpurun_out/test.log
timeout -s KILL 900 python bench.py --steps 24 --warmup 12 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.log | cut -c1-250; tail -5 gpurun_out/bench.err
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
