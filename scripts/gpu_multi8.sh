mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.log 2>&1
echo "exit $?"
python - <<'PY'
import json
for line in open("gpurun_out/bench_8gpu.log"):
    if line.startswith("{"):
        d = json.loads(line); print("value", d["value"], "e2e", d["e2e"]["value"], "n", d["n_gpus"], "ms/step", d["ms_per_step"]); print(json.dumps(d["configs"], indent=1)[:3800])
PY
grep -v "^{" gpurun_out/bench_8gpu.log | tail -5 | cut -c1-300
