#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py tests/test_mel_gpu.py -q -m gpu --timeout 200 > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -60 gpurun_out/test.log
