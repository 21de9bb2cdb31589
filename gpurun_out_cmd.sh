mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/pytest_bwd.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bwd.log; grep -E "Error|error|passed|failed|rc=|assert" gpurun_out/pytest_bwd.log | head -20
timeout 200 python tools/time_train.py > gpurun_out/train_f2.log 2>&1; tail -1 gpurun_out/train_f2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/time_train.py --iters 1 > gpurun_out/ncu_train.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.txt; head -6 gpurun_out/launches_train_summary.txt
