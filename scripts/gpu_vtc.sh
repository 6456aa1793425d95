mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_vtc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_vtc.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_model_vtc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_model_vtc.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vtc.log 2>&1
tail -12 gpurun_out/pytest_vtc.log | cut -c1-200; tail -3 gpurun_out/pytest_model_vtc.log; grep "^{" gpurun_out/bench_vtc.log | tail -1 | cut -c1-250
