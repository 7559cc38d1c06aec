#!/bin/bash
for v in 0 1; do
  echo "== WB_GEMM_2CTA=$v"
  WB_GEMM_2CTA=$v timeout 300 python scripts/gpu_gemm.py time 2>&1 | grep -E "^nq=(128|256|1024)|FAIL" | cut -c1-150
done
WB_GEMM_2CTA=1 timeout 120 python scripts/diag_gemm.py 2>&1 | grep -E "trial" | head -3
