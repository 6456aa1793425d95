mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2>&1
timeout 900 $TR --master-port 29502 scripts/retrieval_eval.py --db 10000 --queries 2000 > gpurun_out/retrieval_${N}gpu.log 2>&1
for f in bench_${N}gpu retrieval_${N}gpu; do echo "== $f"; grep "^{" gpurun_out/$f.log | tail -1 | cut -c1-900; done
