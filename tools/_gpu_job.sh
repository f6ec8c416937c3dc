mkdir -p gpurun_out
for v in tc_s1 tc_p4 tc_p6 tc_p8; do echo "== $v"; COPTER_B200_LIB=tools/variants/lib_$v.so COPTER_B200_POLICY_TC=1 timeout 300 python tools/policy_tc_check.py 2>&1 | head -3; done > gpurun_out/policy_tc_sweep.txt 2>&1; cat gpurun_out/policy_tc_sweep.txt
