# round 2 first check: all GPU tests with the calibrated weights + a bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
tail -30 gpurun_out/pytest_all.log | cut -c1-300; grep "^{" gpurun_out/bench_quick.log | tail -1 | cut -c1-250
