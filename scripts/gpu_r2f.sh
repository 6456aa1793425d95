mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_training_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -25 gpurun_out/pytest_quick.log | cut -c1-300
timeout 300 python scripts/train_profile.py 16 2>&1 | cut -c1-180 | grep -v "^-" | head -34
timeout 300 python scripts/train_profile.py 16 2>&1 | tail -1
