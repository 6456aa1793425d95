# quick loop on the current build: tensor-core kernel + model tests, fp0 timeline, bench with the FP row order on / off
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log | cut -c1-250
timeout 200 python scripts/tc_trace.py fp0 > gpurun_out/trace_fp0.log 2>&1; sed -n 4,9p gpurun_out/trace_fp0.log | cut -c1-220; tail -2 gpurun_out/trace_fp0.log | cut -c1-200
for rep in 1 2; do for o in 0 1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fp-order $o > gpurun_out/bench_ab_$o.log 2>&1
  echo "order $o: $(grep '^{' gpurun_out/bench_ab_$o.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value']), round(d['e2e']['value']), {k: s[k] for k in ('sa0','sa1','sa2','fp2','fp1','fp0')})")"
done; done
