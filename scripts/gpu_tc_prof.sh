mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_tc2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 12 -c 4 -o gpurun_out/prof_tc -f \
    python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/pytest_tc2.log
