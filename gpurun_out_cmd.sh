mkdir -p gpurun_out
echo "=== pytest all gpu"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "AssertionError|passed|failed|^FAILED|rc=" gpurun_out/pytest_gpu.log | head -12
echo "=== timing"
for p in f16x3 bf16x3 bf16; do timeout 120 python tools/time_field.py --precision $p --iters 3 2>&1 | tail -1; done | tee gpurun_out/timing_cg2.log
timeout 120 python tools/time_field.py --precision f16x3 --iters 3 --samples 64 2>&1 | tail -1 | tee -a gpurun_out/timing_cg2.log
echo "=== bench"; timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; head -c 420 gpurun_out/bench_r01.json; echo; tail -c 900 gpurun_out/bench_r01.json; tail -3 gpurun_out/bench_r01.err
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
