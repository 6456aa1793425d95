mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_gpu.py tests/test_pptnet_gpu.py tests/test_reference_python_gpu.py tests/test_pointops_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
tail -30 gpurun_out/pytest_quick.log | cut -c1-400
python - <<'PY'
import json
for line in open("gpurun_out/bench_default.log"):
    if line.startswith("{"):
        d = json.loads(line); print(d["value"], d["e2e"]["value"]); print(json.dumps(d.get("configs"), indent=1)[:3000])
PY
tail -3 gpurun_out/bench_default.log | cut -c1-300
