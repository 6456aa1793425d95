#!/bin/bash
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py -x -q -m gpu --timeout 120 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:afa -c 40 --csv --log-file gpurun_out/afa_launches.csv python scripts/afa_debug.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/afa_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
for r in rows[-8:]:
    print(r[ki][:40], r[gi], r[vi])
PY
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['stage_ms']['afa'])"
