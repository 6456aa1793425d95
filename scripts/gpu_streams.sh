mkdir -p gpurun_out
for cfg in "116 2" "112 2" "120 2" "124 2" "108 2" "116 3" "116 1" "116 2"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tc-ctas $1 --dense-streams $2 > gpurun_out/bench_streams.log 2>&1
  echo "tc_ctas=$1 dense_streams=$2: $(grep '^{' gpurun_out/bench_streams.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))")"
done
