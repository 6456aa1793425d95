mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_pipe.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_pipe.log
for m in stream graph eager; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mode $m > gpurun_out/bench_$m.log 2>&1
done
tail -4 gpurun_out/pytest_pipe.log; for m in stream graph eager; do tail -1 gpurun_out/bench_$m.log | cut -c1-200; done
