#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -k "argmax or adamw" -q -m gpu --timeout 300 > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -80 gpurun_out/test.log
