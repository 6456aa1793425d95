mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 77 --launch-timeout 0 \
  python -m pytest tests/test_pptnet_gpu.py tests/test_training_gpu.py tests/test_losses_retrieval_gpu.py tests/test_pointops_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -x \
  -k "sa_layer_fused and (256-64-2 or 512-16-2 or 128-1-1) or fused_train_bn and (3-7-13 or 2-256-16) or deterministic_backward or split_topk and (7-1000-10 or 5-300-101) or fps_pruned_sampler and 2048" \
  > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/racecheck.log | sort | uniq -c | head -20
