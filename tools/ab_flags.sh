# A/B of experiment switches: TASR_GEMM_FLAGS bits (1 = no L2 prefetch of the next tile's saved tensor, 2 = GEMM grids use
# every SM instead of ceil(tiles / rounds) CTAs), TASR_NO_WGRAD_OVERLAP=1 (weight gradients on the main stream)
run() { env "$@" python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['step_ms_this_batch'],3), round(d['roofline']['gemm_ms_per_step'],3))
"; }
for i in 1 2; do run TASR_GEMM_FLAGS=0; run TASR_GEMM_FLAGS=2; run TASR_GEMM_FLAGS=0 TASR_NO_WGRAD_OVERLAP=1; run TASR_GEMM_FLAGS=2 TASR_NO_WGRAD_OVERLAP=1; done
