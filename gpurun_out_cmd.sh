mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/pytest_bwd.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bwd.log; grep -E "Error|error|passed|failed|rc=|assert" gpurun_out/pytest_bwd.log | head -20
SNB_BWD_SIMT=1 timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/pytest_bwd_simt.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bwd_simt.log; grep -E "passed|failed|rc=" gpurun_out/pytest_bwd_simt.log | head
timeout 200 python tools/time_train.py > gpurun_out/train_tcd.log 2>&1; tail -1 gpurun_out/train_tcd.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/time_train.py --iters 1 > gpurun_out/ncu_train.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.txt; head -9 gpurun_out/launches_train_summary.txt
