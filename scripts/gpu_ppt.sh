mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pptnet_gpu.py tests/test_abi_cpu.py -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_ppt.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ppt.log
timeout 600 python scripts/other_configs.py pptnet > gpurun_out/other_configs.log 2>&1
timeout 300 python scripts/ppt_diag.py > gpurun_out/ppt_diag.log 2>&1
tail -12 gpurun_out/pytest_ppt.log | cut -c1-250; tail -2 gpurun_out/other_configs.log | cut -c1-300; tail -3 gpurun_out/ppt_diag.log | cut -c1-250
