mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01_2gpu.json 2> gpurun_out/bench_r01_2gpu.err; tail -c 1500 gpurun_out/bench_r01_2gpu.json; tail -5 gpurun_out/bench_r01_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_r01_2gpu_ref.json 2> gpurun_out/bench_r01_2gpu_ref.err; tail -c 300 gpurun_out/bench_r01_2gpu_ref.json
