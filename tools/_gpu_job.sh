mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench.json
timeout 600 python bench.py --stream const --no-cpu-baseline > gpurun_out/bench_const.json 2>/dev/null; echo "const exit $?"
timeout 600 python bench.py --stream unif --no-cpu-baseline > gpurun_out/bench_unif.json 2>/dev/null; echo "unif exit $?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref exit $?"
timeout 900 python tools/bench_extra.py > gpurun_out/bench_extra.log 2>&1; echo "extra exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:copter_step_kernel -s 2 -c 2 -o gpurun_out/prof_r1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy -s 2 -c 3 -o gpurun_out/prof_policy3 -f python tools/profile_policy.py > gpurun_out/ncu_policy.log 2>&1; echo "ncu3 exit $?"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/sanitize_racecheck.log
