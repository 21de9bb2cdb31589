mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r01_8gpu.json 2> gpurun_out/bench_r01_8gpu.err; cut -c1-330 gpurun_out/bench_r01_8gpu.json; tail -2 gpurun_out/bench_r01_8gpu.err
