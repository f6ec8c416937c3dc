mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
for v in tc_ro4np tc_ro5np tc_ro6np tc_ro5np80 tc_ro5p80; do echo "== $v"; COPTER_B200_LIB=tools/variants/lib_$v.so timeout 300 python tools/policy_tc_check.py 2>&1 | head -3; done > gpurun_out/r2_sweep_policy_tc_rollout.txt 2>&1; cat gpurun_out/r2_sweep_policy_tc_rollout.txt
