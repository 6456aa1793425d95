mkdir -p gpurun_out
timeout 200 python scripts/fps_bench.py > gpurun_out/fps_bench.log 2>&1; cat gpurun_out/fps_bench.log
timeout 400 python -m pytest tests/test_losses_retrieval_gpu.py tests/test_pointops_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
python - <<'PY'
import json
for line in open("gpurun_out/bench_default.log"):
    if line.startswith("{"):
        d = json.loads(line); print("value", d["value"], "e2e", d["e2e"]["value"]); print(json.dumps(d["configs"]["cfg4_retrieval_10k"])[:400]); print(d["stage_ms"])
PY
