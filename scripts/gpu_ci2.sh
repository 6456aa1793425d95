mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointops_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
timeout 300 python bench.py --steps 20 --warmup 5 --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1
tail -5 gpurun_out/pytest_gpu2.log; tail -3 gpurun_out/bench.log
