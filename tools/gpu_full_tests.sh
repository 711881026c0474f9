#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/test_full.log 2>&1
echo "exit $?" >> gpurun_out/test_full.log
tail -6 gpurun_out/test_full.log
