#!/bin/bash
# kernel + model parity subset, then a short bench (no CPU arm)
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 600 -x -k "${1:-kernels or model or golden}" > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -5 gpurun_out/test.log
timeout -s KILL 900 python bench.py --steps 24 --warmup 12 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'])"; tail -5 gpurun_out/bench.err
