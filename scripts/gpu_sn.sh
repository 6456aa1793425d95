#!/bin/bash
cd "$(dirname "$0")/.."
for v in "--slots 3" "--slots 4" "--slots 5"; do
  echo "$v"
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline $v 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['coalesced']['value']))"
done
