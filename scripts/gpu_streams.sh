mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointops_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider -k fps > gpurun_out/pytest_fps.log 2>&1
tail -2 gpurun_out/pytest_fps.log
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch
from oracle import ops
from patchaugnet_b200 import pointops, _lib as L
rng=np.random.default_rng(0)
for b,n,m in ((5,4096,1024),(32,4096,1024),(3,1024,128),(7,700,100)):
    xyz=rng.uniform(-1,1,(b,n,3)).astype(np.float32); xyz[:, n//2:]=xyz[:, :n-n//2]
    want=ops.furthestsampling(xyz,m)
    L.lib().pab_tune_fps_clouds_per_cta(2)
    got=pointops.furthestsampling(torch.from_numpy(xyz).cuda(), m)
    L.lib().pab_tune_fps_clouds_per_cta(1)
    print('cpc2', b,n,m, bool(torch.equal(got.cpu(), torch.from_numpy(want))))
PY
for cfg in "1" "2" "1" "2"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fps-cpc $cfg > gpurun_out/bench_streams.log 2>&1
  echo "fps_cpc=$cfg: $(grep '^{' gpurun_out/bench_streams.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['stage_ms']['fps0'])")"
done
