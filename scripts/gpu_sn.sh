#!/bin/bash
cd "$(dirname "$0")/.."
export PYTHONDONTWRITEBYTECODE=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 python -m pytest tests/test_mlp_tc_gpu.py tests/test_pptnet_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "pointwise and (128-128 or 512-512 or 64-32) or sa_layer_fused and (256-64-2 or 512-16-2 or 64-1024-2)" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 --launch-timeout 0 python -m pytest tests/test_mlp_tc_gpu.py tests/test_pptnet_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "pointwise and (128-128 or 64-32) or sa_layer_fused and (512-16-2)" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed" | tail -3
for v in "--fps-cpc 1" "--fps-cpc 2"; do
  echo "$v"
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline $v 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['coalesced']['value']))"
done
