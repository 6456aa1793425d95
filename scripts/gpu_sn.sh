#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py -x -q -m gpu --timeout 180 -k "narrow" 2>&1 | tail -2
timeout 100 python scripts/sn_bench.py 2>&1 | head -12
for v in "1,0" "1,3"; do
  echo "PAB_SN=$v"
  PAB_SN=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['coalesced']['value']))"
done
