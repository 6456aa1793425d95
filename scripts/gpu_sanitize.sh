# compute-sanitizer memcheck over the kernels written in round 2 (small shapes)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 \
  python -m pytest tests/test_pptnet_gpu.py tests/test_training_gpu.py tests/test_losses_retrieval_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -x \
  -k "sa_layer_fused and (64-128-1 or 128-256-2 or 64-65-2 or 256-64-2 or 512-16-2 or 128-1-1) or fused_train_bn or deterministic_backward or split_topk and (7-1000-10 or 33-700-128 or 5-300-101) or hard_negatives or emd_matches and 1024-0.05" \
  > gpurun_out/sanitize.log 2>&1
echo "sanitizer exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|Error" gpurun_out/sanitize.log | head -20
