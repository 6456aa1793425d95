# ncu evidence for profiles/: (1) launch list of one bench step (share per kernel), (2) full capture of the top kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointops_gpu.py tests/test_model_gpu.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 12 -c 3 -o gpurun_out/prof_mlp -f \
    python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fps_kernel|knn_kernel|vlad_partial" -c 6 -o gpurun_out/prof_geo -f \
    python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/pytest_gpu3.log; ls -la gpurun_out
