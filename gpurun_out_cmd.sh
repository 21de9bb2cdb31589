mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:field_tc -s 2 -c 1 -o gpurun_out/prof_tc_final2 python tools/time_field.py --precision f16x3 --iters 1 > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:field_tc -s 2 -c 1 -o gpurun_out/prof_tc_final2_bf16 python tools/time_field.py --precision bf16 --iters 1 > gpurun_out/ncu_tc_bf16.log 2>&1; tail -2 gpurun_out/ncu_tc_bf16.log
timeout 300 ncu --set full --clock-control none -k regex:composite_fwd -s 1 -c 1 -o gpurun_out/prof_composite python tools/time_hbm_kernels.py > gpurun_out/ncu_comp.log 2>&1; tail -1 gpurun_out/ncu_comp.log
