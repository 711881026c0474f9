#!/bin/bash
# one ncu --set full pass over the non-GEMM kernels at the bench shapes (one warm-up + one measured launch each)
mkdir -p gpurun_out
BK_ITERS=1 BK_WARMUP=1 python tools/bench_kernels.py > gpurun_out/plain.log 2>&1 &&
BK_ITERS=1 BK_WARMUP=1 ncu --set full --clock-control none --import-source on -k "regex:conv1_|mqa_|attn_|dwconv|gn_fused|bn_silu|bn_bwd|bn_finalize|mel_|colsum" -c 60 -o gpurun_out/prof_kernels -f python tools/bench_kernels.py > gpurun_out/ncu_kernels.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_kernels.log; ls -la gpurun_out/prof_kernels.ncu-rep
