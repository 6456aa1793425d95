mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pptnet_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_ppt.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ppt.log
tail -25 gpurun_out/pytest_ppt.log
