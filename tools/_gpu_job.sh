mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_host_exact.py -m gpu -q -x -k "polic" 2>&1 | tail -5
echo "== default"; timeout 300 python tools/policy_tc_check.py 2>&1 | head -6
echo "== rollout tc=1"; COPTER_B200_POLICY_ROLLOUT_TC=1 timeout 300 python tools/policy_tc_check.py 2>&1 | sed -n 3p
