mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py tests/test_pptnet_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
for cfg in "1 0" "9 0" "1 148" "9 148" "1 132" "1 0"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tc-tune $1 --tc-ctas $2 > gpurun_out/bench_streams.log 2>&1
  echo "tune=$1 tc_ctas=$2: $(grep '^{' gpurun_out/bench_streams.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value']), round(d['e2e']['value']), s['fp0'], s['sa0'], s['sa1'])")"
done
