#!/bin/bash
# diagnostic: segment-kernel variants (particles per thread / occupancy target)
for v in 0 1 2; do
  echo "== HPSIM_SEG_VARIANT=$v"
  HPSIM_SEG_VARIANT=$v python bench.py --steps 3 --warmup 2 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('ms/step', round(d['ms_per_step'], 2), 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'seg us', round(r['avg_launch_us'], 1), 'frac', round(r['frac'], 3), {k: round(v, 2) for k, v in r['by_kernel_ms'].items()})
"
done
