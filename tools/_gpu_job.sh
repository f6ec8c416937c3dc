mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -2
timeout 900 python -m pytest tests/test_gpu_host_exact.py -m gpu -x -q > gpurun_out/pytest_exact.log 2>&1; echo "exact exit $?"; tail -5 gpurun_out/pytest_exact.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_host_exact.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
for k in 1 4 16; do echo "== K=$k"; timeout 600 python tools/sweep.py run --k $k 2>&1 | cut -c1-200; done > gpurun_out/sweep_pair.txt 2>&1; cat gpurun_out/sweep_pair.txt
