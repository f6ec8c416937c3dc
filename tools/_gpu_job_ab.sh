mkdir -p gpurun_out
for v in fresh1 split1 fresh1split1 unroll2; do echo "$v $(COPTER_B200_LIB=tools/variants/lib_$v.so timeout 600 python tools/ab_k.py)" | tee -a gpurun_out/r2_ab_k_loop3.txt; done
echo "default $(timeout 600 python tools/ab_k.py)" | tee -a gpurun_out/r2_ab_k_loop3.txt
