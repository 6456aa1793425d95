mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:"afa_|vlad_finalize|mlp_kernel" -c 12 --csv --log-file gpurun_out/afa_metrics.csv \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_afa.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"afa_fc|afa_att" -c 2 -o gpurun_out/prof_afa -f \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > gpurun_out/ncu_afa2.log 2>&1
ls -la gpurun_out | tail -3
