#!/bin/bash
# sa_narrow_tc kernel inside the stream-mode bench: residency and carve-out variants
cd "$(dirname "$0")/.."
for v in 1,2 1,3 1,4; do
  echo "PAB_SN=$v"
  PAB_SN=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-extras 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['stage_ms']['sa0'])"
done
