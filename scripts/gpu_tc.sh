mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_model_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_model_tc.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_tc.log
tail -5 gpurun_out/pytest_tc.log; tail -3 gpurun_out/pytest_model_tc.log; tail -2 gpurun_out/bench_tc.log | cut -c1-600
