mkdir -p gpurun_out
(for n in 8 48; do for w in seed room; do python tools/grad_error.py $n $w sum; SNB_BWD_SIMT=1 python tools/grad_error.py $n $w sum; done; done; SINNERF_B200_PRECISION=fp32 SNB_BWD_SIMT=1 python tools/grad_error.py 8 seed sum; python tools/grad_error.py 8 seed proj) > gpurun_out/grad_error.log 2>&1
cat gpurun_out/grad_error.log
