mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc|vlad_tc" -c 8 -o gpurun_out/prof_tc4 -f \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_tc4.log 2>&1
tail -3 gpurun_out/pytest_quick.log | cut -c1-300; grep "^{" gpurun_out/bench_quick.log | tail -1 | cut -c1-250
