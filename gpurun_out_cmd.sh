mkdir -p gpurun_out
echo "=== pytest parity (cg2)"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "=== timing"
for d in 0 2; do SNB_TC_DEBUG=$d timeout 120 python tools/time_field.py --precision f16x3 --iters 3 2>&1 | tail -1; done | tee gpurun_out/timing_cg2.log
for p in bf16x3 bf16; do timeout 120 python tools/time_field.py --precision $p --iters 3 2>&1 | tail -1; done | tee -a gpurun_out/timing_cg2.log
timeout 120 python tools/time_field.py --precision f16x3 --iters 3 --samples 64 --sigma-only 2>&1 | tail -1 | tee -a gpurun_out/timing_cg2.log
timeout 120 python tools/trace_field.py f16x3 > gpurun_out/trace_f16x3.log 2>&1; head -43 gpurun_out/trace_f16x3.log | tail -41 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14,$15,$16}' | cut -c1-110; tail -3 gpurun_out/trace_f16x3.log
