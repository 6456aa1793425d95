#!/bin/bash
cd "$(dirname "$0")/.."
for v in "" "--tc-tune 9 --tc-ctas 148" "" "--tc-tune 9 --tc-ctas 148"; do
  echo "$v"
  timeout 200 python bench.py --steps 20 --warmup 5 --no-extras $v 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'])"
done
