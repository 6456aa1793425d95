# round 2 check: all GPU tests + the full default bench line (+ reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>&1
tail -40 gpurun_out/pytest_all.log | cut -c1-400; grep "^{" gpurun_out/bench_default.log | tail -1 | cut -c1-250; tail -3 gpurun_out/bench_default.log | cut -c1-600
