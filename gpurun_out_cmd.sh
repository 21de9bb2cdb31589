mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['train_step'], d['roofline']['frac'], d['roofline']['frac_executed'], d['clocks'])"
wc -l gpurun_out/bench_r01.json
