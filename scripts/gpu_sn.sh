#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py -x -q -m gpu --timeout 120 -k pointwise 2>&1 | tail -8
timeout 900 python -m pytest tests/test_pptnet_gpu.py -x -q -m gpu --timeout 120 2>&1 | tail -8
timeout 300 python scripts/ppt_stages.py 2>&1 | head -24
