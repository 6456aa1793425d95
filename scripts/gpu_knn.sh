mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointops_gpu.py tests/test_refgpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_knn.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_knn.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dense-streams 2 > gpurun_out/bench_ds2.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dense-streams 1 > gpurun_out/bench_ds1.log 2>&1
tail -4 gpurun_out/pytest_knn.log | cut -c1-300; for f in ds2 ds1; do grep "^{" gpurun_out/bench_$f.log | tail -1 | cut -c1-200; done
