#!/bin/bash
# what the driver does at round end, plus the profiles
mkdir -p gpurun_out
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-1800
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout -s KILL 300 python bench.py --model conformer-m --steps 12 --warmup 8 --no-cpu > gpurun_out/bench_m.log 2> gpurun_out/bench_m.err; echo "conformer-m exit $?"; tail -1 gpurun_out/bench_m.log | cut -c1-200
TASR_NO_WGRAD_OVERLAP=1 timeout 300 python tools/kineto_step.py 0 > gpurun_out/kineto.txt 2>&1; echo "kineto exit $?"
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
