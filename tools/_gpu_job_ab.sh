mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "step_c10 $(timeout 600 python tools/ab_k.py)" | tee -a gpurun_out/r2_ab_k_loop5.txt
timeout 300 python tools/bench_extra.py fp64 config2 2>&1 | tail -4
