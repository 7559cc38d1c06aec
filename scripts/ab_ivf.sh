#!/bin/bash
# A/B of IVF gather knobs on one box
for pf in 0 1 0 1; do
  echo "== WB_GATHER_PREFETCH=$pf"
  WB_GATHER_PREFETCH=$pf timeout 300 python scripts/bench_ivf.py 2>&1 | grep search | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['nq'] in (16, 256) and d['nprobe'] in (32, 128): print(d['nq'], d['nprobe'], round(d['scan_ms'], 3), round(d['scan_GBs']))
"
done
