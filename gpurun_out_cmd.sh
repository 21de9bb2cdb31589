mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; cat gpurun_out/bench_r01.json | cut -c1-400
python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_r01_bf16.json 2> gpurun_out/bench_r01_bf16.err; cat gpurun_out/bench_r01_bf16.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f16x3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:field_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/prof_tc_v11 -f python tools/time_field.py --precision f16x3 --iters 1 > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:field_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/prof_tc_v11_bf16 -f python tools/time_field.py --precision bf16 --iters 1 > gpurun_out/ncu_tc_bf16.log 2>&1
for pr in f16x3 bf16x3 bf16; do timeout 120 python tools/time_field.py --precision $pr --iters 5; done > gpurun_out/timing_v11.log 2>&1
timeout 120 python tools/time_field.py --precision f16x3 --samples 64 --sigma-only --iters 5 >> gpurun_out/timing_v11.log 2>&1
