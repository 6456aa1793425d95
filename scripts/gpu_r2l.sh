mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_pptnet_gpu.py tests/test_reference_python_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x 2>&1 | tail -8 | cut -c1-250
timeout 200 python scripts/ppt_stages.py 64 2>&1 | head -22
