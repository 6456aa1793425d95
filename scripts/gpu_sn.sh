#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_pptnet_gpu.py -x -q -m gpu --timeout 180 -k "stream or bf16" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['configs']['cfg3_pptnet_b64'], indent=1)[:1800])"
