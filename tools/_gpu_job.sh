mkdir -p gpurun_out
./tools/microbench/ffma2_rates > gpurun_out/ffma2_rates.txt 2>&1; grep -E "3reg|2reg|ILP 8" gpurun_out/ffma2_rates.txt | tail -30
timeout 900 python -m pytest tests/test_gpu_host_exact.py -m gpu -q > gpurun_out/pytest_exact.log 2>&1; echo "exact exit $?"; tail -5 gpurun_out/pytest_exact.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:copter_step -s 2 -c 1 -o gpurun_out/prof_k16_scalar -f python tools/profile_k.py 16 > gpurun_out/ncu_k16_scalar.log 2>&1; echo "ncu scalar exit $?"
COPTER_B200_PAIR_MIN_K=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:copter_step -s 2 -c 1 -o gpurun_out/prof_k16_pair -f python tools/profile_k.py 16 > gpurun_out/ncu_k16_pair.log 2>&1; echo "ncu pair exit $?"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_try.json 2> gpurun_out/bench_try.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/bench_try.json; tail -5 gpurun_out/bench_try.err
