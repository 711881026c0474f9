#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -25 gpurun_out/test.log
