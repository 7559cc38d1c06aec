#!/bin/bash
timeout 200 python scripts/diag_gemm.py 2>&1 | grep trial
timeout 300 python scripts/gpu_gemm.py time 2>&1 | cut -c1-170
