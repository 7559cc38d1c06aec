#!/bin/bash
for c in 0 1; do
  echo "== WB_IVF_CONTIGUOUS=$c"
  WB_IVF_CONTIGUOUS=$c timeout 300 python scripts/bench_ivf.py 2>&1 | grep search | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['nq'], d['nprobe'], 'scan_ms', round(d['scan_ms'], 3), 'GB/s', round(d['scan_GBs']), 'call_ms', round(d['call_ms'],3), 'recall', d['recall_vs_flat'])
"
done
