#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_pointops_gpu.py -x -q -m gpu --timeout 180 2>&1 | tail -3
for v in "" ""; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline $v 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['coalesced']['value']))"
done
