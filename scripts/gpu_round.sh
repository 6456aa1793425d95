# full round check: every GPU test, the default bench line, the reference arm, the ncu launch list and full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc|vlad_tc|knn_pruned32|fps_kernel|three_nn_pruned" -c 15 -o gpurun_out/r01_top -f \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_top.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -4 gpurun_out/pytest_all.log | cut -c1-300; grep "^{" gpurun_out/bench_default.log | tail -1 | cut -c1-300; tail -2 gpurun_out/smoke.log
timeout 300 python scripts/other_configs.py pptnet > gpurun_out/other_configs.log 2>&1; tail -2 gpurun_out/other_configs.log | cut -c1-300
timeout 100 python scripts/tc_trace.py fp0 > gpurun_out/trace_fp0.log 2>&1; timeout 100 python scripts/vlad_trace.py vlad2 > gpurun_out/trace_vlad2.log 2>&1
