#!/bin/bash
mkdir -p gpurun_out
python tools/prof_gemm.py once 0.1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 6 -c 6 -o gpurun_out/prof_gemm2 -f python tools/prof_gemm.py once 0.1 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
