mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1
python - <<'PY'
import json
for line in open("gpurun_out/bench_default.log"):
    if line.startswith("{"):
        d = json.loads(line); print("value", d["value"], "e2e", d["e2e"]["value"])
        c = d["configs"]
        print("cfg3", json.dumps(c["cfg3_pptnet_b64"].get("stock_gpu"))[:300], c["cfg3_pptnet_b64"]["f32"])
        print("cfg4", json.dumps(c["cfg4_retrieval_10k"])[:300])
        print("cfg5", c["cfg5_train_step"].get("ms_per_step"), json.dumps(c["cfg5_train_step"].get("stock_gpu"))[:300])
PY
tail -2 gpurun_out/bench_default.log | cut -c1-200
