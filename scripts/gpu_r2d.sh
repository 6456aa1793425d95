mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_pptnet_gpu.py tests/test_mlp_tc_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 60 -x > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
timeout 300 python scripts/ppt_stages.py 64 > gpurun_out/ppt_stages.log 2>&1
tail -40 gpurun_out/pytest_quick.log | cut -c1-300; head -30 gpurun_out/ppt_stages.log
