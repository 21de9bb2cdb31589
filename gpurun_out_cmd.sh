mkdir -p gpurun_out
echo "=== pytest (cta_group 2)"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "=== timing"
for d in 0 2 4 6; do SNB_TC_DEBUG=$d timeout 120 python tools/time_field.py --precision f16x3 --iters 3 2>&1 | tail -1; done | tee gpurun_out/timing_cg2.log
for p in bf16x3 bf16; do timeout 120 python tools/time_field.py --precision $p --iters 3 2>&1 | tail -1; done | tee -a gpurun_out/timing_cg2.log
