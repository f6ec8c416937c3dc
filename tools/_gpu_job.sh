mkdir -p gpurun_out
echo "groundff $(timeout 600 python tools/ab_k.py)" | tee -a gpurun_out/r2_ab_k_loop2.txt
echo "noff $(COPTER_B200_LIB=tools/variants/lib_noff.so timeout 600 python tools/ab_k.py)" | tee -a gpurun_out/r2_ab_k_loop2.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
