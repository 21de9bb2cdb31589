mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/pytest_bwd.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bwd.log; grep -E "Error|error|passed|failed|rc=" gpurun_out/pytest_bwd.log | head -20
timeout 200 python tools/time_train.py > gpurun_out/train_sep.log 2>&1; tail -1 gpurun_out/train_sep.log
timeout 200 python tools/time_train.py --multi > gpurun_out/train_multi.log 2>&1; tail -1 gpurun_out/train_multi.log
