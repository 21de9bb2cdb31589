mkdir -p gpurun_out
echo "=== probe" ; timeout 120 ./probes/umma_probe > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log; cat gpurun_out/probe.log
echo "=== pytest fp32"; SINNERF_B200_TEST_MODES=fp32 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_fp32.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fp32.log; tail -15 gpurun_out/pytest_fp32.log
echo "=== pytest tc mlp"; SINNERF_B200_TEST_MODES=f16x3,bf16x3,bf16 timeout 300 python -m pytest tests -m gpu -q -k "mlp_forward" > gpurun_out/pytest_tc_mlp.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc_mlp.log; tail -25 gpurun_out/pytest_tc_mlp.log
echo "=== pytest tc all"; SINNERF_B200_TEST_MODES=f16x3,bf16x3,bf16 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc.log; tail -25 gpurun_out/pytest_tc.log
echo "=== bench f16x3"; timeout 300 python bench.py --steps 5 --warmup 3 --precision f16x3 > gpurun_out/bench_f16x3.json 2> gpurun_out/bench_f16x3.err; tail -c 2500 gpurun_out/bench_f16x3.json; tail -3 gpurun_out/bench_f16x3.err
