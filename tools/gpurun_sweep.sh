#!/bin/bash
for c in "$@"; do
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --chunk-pairs $c 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk', $c, round(d['value']), round(d['ms_per_step'],1), round(d['roofline']['kernel_ms_per_step'],1))"
done
