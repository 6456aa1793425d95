mkdir -p gpurun_out
timeout 300 python scripts/ppt_stages.py 64 > gpurun_out/ppt_stages.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --fps-cpc 2 > gpurun_out/bench_cpc2.log 2>&1
timeout 600 python -m pytest tests/test_refgpu.py tests/test_pointops_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
cat gpurun_out/ppt_stages.log; grep "^{" gpurun_out/bench_cpc2.log | tail -1 | cut -c1-200; tail -5 gpurun_out/pytest_quick.log
