# A/B of experiment switches: TASR_GEMM_FLAGS bits (1 = no L2 prefetch of the next tile's saved tensor, 2 = GEMM grids use
# every SM instead of ceil(tiles / rounds) CTAs, 4 = weight-gradient kernels with the deep 4 / 6-stage ring, 32 = narrow
# weight-gradient tiles with 3 stages, 64 = 128-wide tiles for the K >= 2048 dgrad),
# TASR_NO_WGRAD_OVERLAP=1 (weight gradients on the main stream).  Run-to-run noise of the step time is about +-0.5 %:
# repeat every setting (AB_REPS, default 3).
run() { env "$@" python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['step_ms_this_batch'],3), round(d['roofline']['gemm_ms_per_step'],3))
"; }
for i in $(seq ${AB_REPS:-3}); do for f in ${AB_FLAGS:-0 4}; do run TASR_GEMM_FLAGS=$f; done; done
