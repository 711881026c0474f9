#!/bin/bash
# ncu --set full of the kernels matching $1 (regex) inside one profiled training step; report -> gpurun_out/prof_$2.ncu-rep
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$1" -c ${3:-4} -o gpurun_out/prof_$2 -f python tools/profile_step.py > gpurun_out/ncu_$2.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_$2.log; ls -la gpurun_out/prof_$2.ncu-rep
