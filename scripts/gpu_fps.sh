mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointops_gpu.py tests/test_refgpu.py tests/test_pptnet_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_fps.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fps.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
tail -12 gpurun_out/pytest_fps.log | cut -c1-250; grep "^{" gpurun_out/bench_quick.log | tail -1 | cut -c1-200
