#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_augment.py -q -m gpu --timeout 600 -x > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -40 gpurun_out/test.log
