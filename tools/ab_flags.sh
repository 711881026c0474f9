for i in 1 2; do for f in 0 1; do TASR_GEMM_FLAGS=$f python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FLAGS',$f, round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['step_ms_this_batch'],3), round(d['roofline']['gemm_ms_per_step'],3))
"; done; done
