#!/bin/bash
# N-GPU measurement set for profiles/: data-parallel parity check, bench.py, fault sweeps + inference.
# usage: NG=8 bash tools/gpu_scale.sh
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
if [ "$NG" -gt 1 ]; then
  timeout 300 $TR --master-port 29501 tools/dp_check.py > gpurun_out/dp_check_$NG.log 2>&1; echo "dp_check rc=$?"; tail -2 gpurun_out/dp_check_$NG.log
  timeout 400 $TR --master-port 29502 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/bench_n$NG.log 2>&1; echo "bench rc=$?"
  timeout 400 $TR --master-port 29503 tools/bench_sweep.py --trials 64 --cpu-trials 0 > gpurun_out/sweep_n$NG.log 2>&1; echo "sweep rc=$?"
else
  timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"
  timeout 400 python tools/bench_sweep.py --trials 64 --cpu-trials 1 > gpurun_out/sweep_n1.log 2>&1; echo "sweep rc=$?"
fi
tail -1 gpurun_out/bench_n$NG.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'clocks', d['clocks'])"
tail -1 gpurun_out/sweep_n$NG.log | cut -c1-900
