#!/bin/bash
# Multi-GPU checks: N-GPU == 1-GPU == reference, then the scaling bench lines.  Usage: bash tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29611 tools/dp_check.py > gpurun_out/dp_check_$N.log 2>&1
echo "dp_check rc=$? $(tail -n 2 gpurun_out/dp_check_$N.log | tr '\n' ' ' | cut -c1-300)"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --profile > gpurun_out/bench_n1.log 2>&1
echo "bench1 rc=$? $(tail -n 1 gpurun_out/bench_n1.log | cut -c1-200)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1
echo "bench$N rc=$? $(tail -n 1 gpurun_out/bench_n$N.log | cut -c1-400)"
