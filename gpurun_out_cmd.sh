mkdir -p gpurun_out
echo "=== pytest all modes (cluster 2)"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
echo "=== pytest cluster 4 (mlp + render)"; SNB_TC_CLUSTER=4 timeout 600 python -m pytest tests -m gpu -q -x -k "mlp_forward or render_rays_vs" > gpurun_out/pytest_gpu_c4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_c4.log; tail -3 gpurun_out/pytest_gpu_c4.log
echo "=== timing"
for c in 1 2 4; do for d in 0 6; do SNB_TC_CLUSTER=$c SNB_TC_DEBUG=$d timeout 120 python tools/time_field.py --precision f16x3 --iters 3 2>&1 | tail -1 | sed "s/^/cluster=$c /"; done; done | tee gpurun_out/timing_cluster.log
for c in 2 4; do SNB_TC_CLUSTER=$c timeout 120 python tools/time_field.py --precision bf16 --iters 3 2>&1 | tail -1 | sed "s/^/cluster=$c /"; done | tee -a gpurun_out/timing_cluster.log
for c in 2 4; do SNB_TC_CLUSTER=$c SNB_TC_DEBUG=2 timeout 120 python tools/time_field.py --precision f16x3 --iters 3 2>&1 | tail -1 | sed "s/^/cluster=$c /"; done | tee -a gpurun_out/timing_cluster.log
