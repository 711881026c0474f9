#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --model conformer-m --steps 12 --warmup 6 --no-cpu > gpurun_out/bench_m.log 2> gpurun_out/bench_m.err; echo "bench exit $?"; tail -1 gpurun_out/bench_m.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])"; tail -5 gpurun_out/bench_m.err
