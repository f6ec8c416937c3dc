# usage: bash tools/_gpu_job_multi.sh N     (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
PORT=29611
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv | tail -$N
lscpu | grep -E "^CPU\(s\)|NUMA node|Socket|Model name" | head -6
timeout 900 $TR bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "weak exit $?"; cut -c1-260 gpurun_out/r2_bench_${N}gpu.json; tail -2 gpurun_out/bench_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --global-envs 16777216 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/r2_bench_strong_${N}gpu.json 2> gpurun_out/bench_strong_${N}gpu.err; echo "strong exit $?"; cut -c1-260 gpurun_out/r2_bench_strong_${N}gpu.json; tail -2 gpurun_out/bench_strong_${N}gpu.err
timeout 600 $TR tools/pcie_probe.py > gpurun_out/r2_pcie_probe_${N}gpu.json 2>/dev/null; echo "pcie exit $?"; cut -c1-600 gpurun_out/r2_pcie_probe_${N}gpu.json
