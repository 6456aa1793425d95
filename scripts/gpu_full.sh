# every GPU test (per-test timeout) + the default bench line + the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 180 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1
tail -30 gpurun_out/pytest_all.log | cut -c1-300
python - <<'PY'
import json
for line in open("gpurun_out/bench_default.log"):
    if line.startswith("{"):
        d = json.loads(line); print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
        print("stock", d.get("stock_gpu")); print(json.dumps(d.get("configs"), indent=1)[:2500])
PY
tail -2 gpurun_out/bench_default.log | cut -c1-200
