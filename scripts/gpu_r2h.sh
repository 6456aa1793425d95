mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_pptnet_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 2>&1 | tail -8 | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
python - <<'PY'
import json
for line in open("gpurun_out/bench_default.log"):
    if line.startswith("{"):
        d = json.loads(line); print("value", d["value"], "e2e", d["e2e"]["value"]); print(json.dumps(d["configs"]["cfg3_pptnet_b64"], indent=1)[:1200]); print(json.dumps(d["configs"]["cfg4_retrieval_10k"])[:300])
PY
tail -2 gpurun_out/bench_default.log | cut -c1-300
