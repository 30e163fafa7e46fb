#!/bin/bash
# N-GPU runs of both workloads, launched exactly like the driver does (torchrun, one rank per GPU)
cd "$(dirname "$0")/.."
N=${1:-2}; tag=${2:-r01_d}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 900 bash -c "$(declare -f run); N=$N; run 29511 --steps 2 --warmup 3" > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; echo "sample exit $?"; cut -c1-260 gpurun_out/bench_${N}gpu_$tag.json
timeout 900 bash -c "$(declare -f run); N=$N; run 29512 --workload train --steps 5 --warmup 3" > gpurun_out/bench_train_${N}gpu_$tag.json 2> gpurun_out/bench_train_${N}gpu_$tag.err; echo "train exit $?"; cut -c1-260 gpurun_out/bench_train_${N}gpu_$tag.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/dp_check.py > gpurun_out/dp_check_${N}gpu_$tag.log 2>&1; echo "dp_check exit $?"; tail -3 gpurun_out/dp_check_${N}gpu_$tag.log
