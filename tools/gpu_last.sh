#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout -s KILL 600 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/test_full.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/test_full.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
BK_ITERS=1 BK_WARMUP=1 timeout 300 ncu --set full --clock-control none -k "regex:conv1_tc|mqa_fwd|mqa_bwd|gn_fused|dwconv|mel_logpower" -c 24 -o gpurun_out/prof_final_kernels -f python tools/bench_kernels.py mel conv1 groupnorm dwconv mqa > gpurun_out/ncu_final.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_final_kernels.ncu-rep
