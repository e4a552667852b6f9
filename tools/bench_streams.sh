#!/bin/bash
# Developer tool: bench.py at several --streams values (one line each).
for s in "$@"; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams $s 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams', $s, 'value', d['value'], 'e2e', d['e2e']['value'])"
done
