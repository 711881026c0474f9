#!/bin/bash
# A/B of the GEMM shapes: current library vs tools/_ab_base.so (device times from CUPTI)
echo "== new"; timeout 120 python tools/prof_gemm.py all 0.1 2>&1 | grep -v -i warn
cp tools/_ab_base.so turkish_asr_model_b200/libtasr_kernels.so
echo "== base"; timeout 120 python tools/prof_gemm.py all 0.1 2>&1 | grep -v -i warn
