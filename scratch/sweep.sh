#!/bin/bash
# tuning sweep: window bits and chunk granularity on the 2^20 prove
for c in 16 15 14; do for k in 4 2 8; do
  echo "== ZA_MSM_C=$c CHUNKS=$k"
  ZA_MSM_C=$c ZA_MSM_CHUNKS_PER_SLOT=$k python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],2), round(d['e2e']['value'],2), d['kernel_ms_per_step'])"
done; done
