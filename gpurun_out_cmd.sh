mkdir -p gpurun_out
echo "=== pytest all gpu"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "AssertionError|Error|passed|failed|^FAILED|rc=" gpurun_out/pytest_gpu.log | head -12
