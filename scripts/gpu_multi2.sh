mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
echo "exit $?"
python - <<'PY'
import json
for line in open("gpurun_out/bench_2gpu.log"):
    if line.startswith("{"):
        d = json.loads(line); print("value", d["value"], "e2e", d["e2e"]["value"], "n", d["n_gpus"]); print(json.dumps(d["configs"], indent=1)[:3500])
PY
tail -5 gpurun_out/bench_2gpu.log | cut -c1-400
