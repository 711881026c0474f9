#!/bin/bash
# A/B of an environment switch ($1 = variable name): model tests once, then bench with VAR=1 and VAR=0
mkdir -p gpurun_out
VAR=${1:-TASR_CONV_MC}
timeout -s KILL 300 python -m pytest tests -q -m gpu --timeout 120 -x -k "conv or model or golden" > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -3 gpurun_out/test.log
for v in 1 0; do
env $VAR=$v timeout -s KILL 300 python bench.py --steps 24 --warmup 12 --no-cpu > gpurun_out/bench_ab$v.log 2> gpurun_out/bench_ab$v.err; echo "$VAR=$v exit $?"; tail -1 gpurun_out/bench_ab$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -3 gpurun_out/bench_ab$v.err
done
