mkdir -p gpurun_out
echo "numer2 $(timeout 600 python tools/ab_k.py)" | tee -a gpurun_out/r2_ab_k_loop2.txt
bash tools/_ncu_capture.sh step_kernel_k16 copter_step 2 1 8388608 -- python tools/profile_k.py 16
head -16 gpurun_out/r2_step_kernel_k16_sass_mix.txt; grep "duration\|issue_active\|registers_per" gpurun_out/r2_step_kernel_k16_ncu_full.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
