mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fps_|knn_|three_nn|mlp_|vlad_|attn_|gather_rows|gemm|gemv|elementwise|softmax|reduce" -c 260 --csv --log-file gpurun_out/ppt_launches.csv \
    python scripts/other_configs.py pptnet > gpurun_out/ncu_ppt.log 2>&1
tail -2 gpurun_out/ncu_ppt.log | cut -c1-200
