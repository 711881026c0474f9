#!/bin/bash
mkdir -p gpurun_out
python tools/prof_gemm.py all 0.1 > gpurun_out/gemm_shapes.log 2>&1; cat gpurun_out/gemm_shapes.log
python tools/prof_gemm.py all 0.0 2>&1 | tee gpurun_out/gemm_shapes_nodrop.log
python tools/prof_gemm.py ff1 0.1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_ff1 -f python tools/prof_gemm.py ff1 0.1 > gpurun_out/ncu_ff1.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
