mkdir -p gpurun_out
echo "=== pytest all gpu"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "AssertionError|passed|failed|^FAILED|rc=" gpurun_out/pytest_gpu.log | head -8
timeout 300 python tools/time_hbm_kernels.py > gpurun_out/hbm_kernels.log 2>&1; cat gpurun_out/hbm_kernels.log
echo "=== bench"; timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; head -c 330 gpurun_out/bench_r01.json; echo; tail -3 gpurun_out/bench_r01.err
