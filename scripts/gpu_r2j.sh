mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_model_gpu.py tests/test_input_pipeline_gpu.py tests/test_losses_retrieval_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x 2>&1 | tail -6 | cut -c1-250
for sg in 1 0 1 0; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --stream-graphs $sg > gpurun_out/bench_sg$sg.log 2>&1
python - <<PY
import json
for line in open("gpurun_out/bench_sg$sg.log"):
    if line.startswith("{"):
        d = json.loads(line); print("stream-graphs $sg value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
PY
done
tail -3 gpurun_out/bench_sg1.log | cut -c1-300
