mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi" 2>&1 | tail -2
timeout 200 python tools/time_train.py > gpurun_out/train_sep.log 2>&1; tail -1 gpurun_out/train_sep.log
timeout 200 python tools/time_train.py --multi > gpurun_out/train_multi.log 2>&1; tail -1 gpurun_out/train_multi.log
