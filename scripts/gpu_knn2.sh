mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointops_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_knn.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_knn.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"knn_pruned|knn_index|three_nn_pruned" -c 5 --csv --log-file gpurun_out/knn_metrics.csv \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_knn.log 2>&1
tail -12 gpurun_out/pytest_knn.log | cut -c1-300; grep "^{" gpurun_out/bench_quick.log | tail -1 | cut -c1-250
