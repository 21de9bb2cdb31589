mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q -x -k "additive" 2>&1 | tail -2
timeout 200 python tools/time_configs.py > gpurun_out/time_configs.log 2>&1; cat gpurun_out/time_configs.log
