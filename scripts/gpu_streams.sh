mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointops_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider -k fps > gpurun_out/pytest_fps.log 2>&1
tail -2 gpurun_out/pytest_fps.log
for cfg in 0 256 512 0 256; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fps-threads $cfg > gpurun_out/bench_streams.log 2>&1
  echo "fps_threads=$cfg: $(grep '^{' gpurun_out/bench_streams.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['stage_ms']['fps0'], d['stage_ms']['fps1'])")"
done
