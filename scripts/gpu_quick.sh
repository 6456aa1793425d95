# quick loop: model-level GPU tests + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py tests/test_pptnet_gpu.py tests/test_losses_retrieval_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
tail -4 gpurun_out/pytest_quick.log | cut -c1-300; grep "^{" gpurun_out/bench_quick.log | tail -1 | cut -c1-250
