mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointops_gpu.py tests/test_model_gpu.py tests/test_losses_retrieval_gpu.py tests/test_refgpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
timeout 300 python scripts/fps_bench.py > gpurun_out/fps_bench.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_quick.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --fps-pruned 0 --fps-cpc 1 > gpurun_out/bench_quick_old.log 2>&1
tail -25 gpurun_out/pytest_quick.log | cut -c1-300; cat gpurun_out/fps_bench.log; for f in gpurun_out/bench_quick.log gpurun_out/bench_quick_old.log; do grep "^{" $f | tail -1 | cut -c1-200; done
