#!/bin/bash
# A/B: bench with and without the side-stream weight-gradient overlap
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -q -m gpu --timeout 600 -x -k "model or golden" > gpurun_out/test.log 2>&1
echo "exit $?" >> gpurun_out/test.log
tail -3 gpurun_out/test.log
for v in 1 0; do
TASR_CAP_PRIO=$v timeout -s KILL 300 python bench.py --steps 24 --warmup 12 --no-cpu > gpurun_out/bench_ov$v.log 2> gpurun_out/bench_ov$v.err; echo "cap_prio=$v exit $?"; tail -1 gpurun_out/bench_ov$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -3 gpurun_out/bench_ov$v.err
done
