# round-2 ncu evidence: per-kernel metrics of one whole eager forward (29 launches), full-set capture of the top kernels,
# metrics of the kernels outside the bench step (attention, retrieval top-k, training kernels)
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size
ALL="fps_kernel|gather_rows_kernel|knn_index_kernel|knn_pruned|knn_kernel|three_nn|mlp_tc_kernel|sa_narrow_tc_kernel|vlad_tc_kernel|vlad_finalize_kernel|afa_"
timeout 900 ncu --metrics $M --clock-control none -k regex:"$ALL" -s 58 -c 29 --csv --log-file gpurun_out/r02_step_metrics.csv \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-extras --repeats 1 > gpurun_out/ncu_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_kernel|sa_narrow_tc_kernel|vlad_tc_kernel|afa_fc_tc_kernel|afa_att_tc_kernel|fps_kernel|knn_pruned32|three_nn_pruned" -s 36 -c 18 -o gpurun_out/r02_top -f \
    python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline --no-extras --repeats 1 > gpurun_out/ncu_top.log 2>&1
ncu -i gpurun_out/r02_top.ncu-rep --page raw --csv --metrics gpu__time_duration.sum | cut -d, -f5,8,9,12 | head -20
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_step_metrics.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
seen=collections.OrderedDict()
for r in rows[h+1:]:
    if r and r[0].isdigit(): seen.setdefault(r[0], r[4].split('(')[0].replace("void ","")[:40])
print(list(seen.values()))
PY
