#!/bin/bash
for c in "0 0" "0 1" "1 0" "1 1"; do
  set -- $c
  echo "== WB_GEMM_2CTA=$1 WB_GEMM_PREFETCH=$2"
  WB_GEMM_2CTA=$1 WB_GEMM_PREFETCH=$2 timeout 300 python scripts/gpu_gemm.py time 2>&1 | grep -E "^nq=(16|64|128) |^nq=(16|64|128):|FAIL" | cut -c1-110
done
