#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_mlp_tc_gpu.py tests/test_pptnet_gpu.py -x -q -m gpu --timeout 180 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['coalesced']['value']))"
