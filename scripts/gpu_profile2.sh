# ncu evidence of the tensor-core path for profiles/: (1) launch list of eager bench steps, (2) full captures of the top kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --steps 2 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_bench_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc|vlad_tc" -c 7 -o gpurun_out/prof_tc2 -f \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_tc2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fps_kernel|knn_kernel|three_nn|afa_|mlp_kernel" -c 12 -o gpurun_out/prof_geo2 -f \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_geo2.log 2>&1
ls -la gpurun_out | tail -8
