#!/bin/bash
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
echo "exit $?"; tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
