mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pptnet_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_ppt.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ppt.log
timeout 600 python scripts/retrieval_eval.py --db 2000 --queries 400 > gpurun_out/retrieval_1gpu.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2.log 2>&1
tail -4 gpurun_out/pytest_ppt.log; tail -2 gpurun_out/retrieval_1gpu.log | cut -c1-900; tail -1 gpurun_out/bench_r2.log | cut -c1-300
