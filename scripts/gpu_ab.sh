mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
for rep in 1 2; do for d in ab_1fc22de .; do
  (cd $d && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $GRAFT_REPO_ROOT/gpurun_out/ab${rep}_$(basename $d).log 2>&1)
done; done
tail -3 gpurun_out/pytest_quick.log
for f in gpurun_out/ab1_*.log gpurun_out/ab2_*.log; do echo $f; grep "^{" $f | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print(round(d['value']), {k:s[k] for k in ('fp0','sa0','sa1','sa2','fp1','fp2','vlad2')})"; done
