mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py tests/test_pptnet_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x 2>&1 | tail -6 | cut -c1-250
for tune in 1 33 1 33; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --tc-tune $tune > gpurun_out/bench_tune$tune.log 2>&1
python - <<PY
import json
for line in open("gpurun_out/bench_tune$tune.log"):
    if line.startswith("{"):
        d = json.loads(line); s = d["stage_ms"]; print("tune $tune value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "sa0", s["sa0"], "sa1", s["sa1"], "sa2", s["sa2"], "fp0", s["fp0"])
PY
done
for i in 1 2 3; do timeout 120 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 60 -x 2>&1 | tail -1; done
