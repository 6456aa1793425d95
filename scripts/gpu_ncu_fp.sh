# ncu full capture of the six fused-MLP launches of one eager forward, FP row order on and off
mkdir -p gpurun_out
for o in 1 0; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc" --launch-skip 18 -c 6 -o gpurun_out/r01_mlp_order$o -f \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --fp-order $o > gpurun_out/ncu_mlp_order$o.log 2>&1
ncu -i gpurun_out/r01_mlp_order$o.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,sm__inst_executed_pipe_tmem.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/r01_mlp_order$o.csv 2>&1
done
cat gpurun_out/r01_mlp_order1.csv | cut -c1-400 | tail -8
