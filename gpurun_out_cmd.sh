mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "other_training" > gpurun_out/pytest_bwd2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bwd2.log; grep -E "Error|error|passed|failed|rc=|assert" gpurun_out/pytest_bwd2.log | head -20
