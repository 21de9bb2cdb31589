mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for pr in f16x3 bf16; do timeout 120 python tools/time_field.py --precision $pr --iters 5; done > gpurun_out/timing_v12b.log 2>&1
cat gpurun_out/timing_v12b.log
timeout 120 python tools/trace_field.py f16x3 > gpurun_out/trace_f16x3_v12b.log 2>&1
git stash -q 2>/dev/null; true
