mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_host_exact.py -m gpu -q -k "polic" 2>&1 | tail -15
echo "== default"; timeout 300 python tools/policy_tc_check.py 2>&1 | head -6
