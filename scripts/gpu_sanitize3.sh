# compute-sanitizer over the kernels added late in round 2: small-CTA SA0 kernel (sa_narrow_tc.cu), tensor-core AFA head and gated
# fc (afa_tc.cu), self-resetting tile counters of mlp_tc.cu (dynamic tiles), SA gather loader lane mapping
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
SEL="sa_narrow_kernel and (spec1 or spec2 or spec3 or spec4 or spec5) or afa_head and (5-256 or 3-64 or 130-128 or 1-192) or scheduling_options or sa_module_tensor_core and (spec1 or spec4)"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 \
  python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -x -k "$SEL" > gpurun_out/memcheck3.log 2>&1
echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/memcheck3.log | sort | uniq -c | head -20
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 77 --launch-timeout 0 \
  python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q -p no:cacheprovider --timeout 1200 -x -k "sa_narrow_kernel and (spec1 or spec2 or spec5) or afa_head and (5-256 or 3-64)" > gpurun_out/racecheck3.log 2>&1
echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/racecheck3.log | sort | uniq -c | head -20
