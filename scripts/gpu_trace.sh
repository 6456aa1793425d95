mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
timeout 300 python scripts/tc_trace.py fp0 > gpurun_out/trace_fp0.log 2>&1
timeout 300 python scripts/tc_trace.py sa0 > gpurun_out/trace_sa0.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
tail -3 gpurun_out/pytest_quick.log | cut -c1-200; sed -n 7,12p gpurun_out/trace_fp0.log; grep "^{" gpurun_out/bench_quick.log | tail -1 | cut -c1-200
