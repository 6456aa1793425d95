mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
timeout 600 $TR --master-port 29502 scripts/retrieval_eval.py --db 4000 --queries 800 > gpurun_out/retrieval_2gpu.log 2>&1
timeout 600 python scripts/retrieval_eval.py --db 4000 --queries 800 > gpurun_out/retrieval_1gpu.log 2>&1
timeout 600 python scripts/stock_gpu_baseline.py > gpurun_out/stock_gpu.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.log 2>&1
for f in bench_2gpu retrieval_2gpu retrieval_1gpu stock_gpu bench_1gpu; do echo "== $f"; grep "^{" gpurun_out/$f.log | tail -1 | cut -c1-700; done
