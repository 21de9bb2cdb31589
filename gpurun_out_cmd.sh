mkdir -p gpurun_out
for cfg in "48 64" "48 128" "333 64"; do set -- $cfg; echo "== N=$1 S=$2"; DBG_N=$1 DBG_S=$2 timeout 300 python tools/debug_bwd.py 2>&1 | awk '{print $1, $NF}' | sort -k2 -g | tail -4; done > gpurun_out/debug_bwd.log 2>&1; cat gpurun_out/debug_bwd.log
echo "=== pytest backward"; timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q > gpurun_out/pytest_bwd.log 2>&1; grep -E "AssertionError: \(|passed|failed" gpurun_out/pytest_bwd.log | head
