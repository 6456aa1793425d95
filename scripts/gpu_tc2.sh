mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_tc3.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc2.log 2>&1
timeout 300 python scripts/microbench.py > gpurun_out/microbench.log 2>&1
tail -3 gpurun_out/pytest_tc3.log; cat gpurun_out/microbench.log | tail -20
