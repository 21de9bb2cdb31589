mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01_2gpu.json 2> gpurun_out/bench_r01_2gpu.err; cut -c1-330 gpurun_out/bench_r01_2gpu.json; tail -2 gpurun_out/bench_r01_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r01_2gpu_ref.json 2> gpurun_out/bench_r01_2gpu_ref.err; cut -c1-200 gpurun_out/bench_r01_2gpu_ref.json
timeout 300 python -m pytest tests -m gpu -q -x -k "shard or distributed or multi" 2>&1 | tail -2
