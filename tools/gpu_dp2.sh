#!/bin/bash
# 2-GPU checks of the data-parallel path: parity against the golden full-batch steps, bench, timeline.
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then
  PT="python -m pytest -q -m gpu -p no:cacheprovider -x"
  timeout 300 $PT tests/test_kernels_gpu.py -k "gemm or variants or precision" > gpurun_out/k_gemm.log 2>&1; echo "k_gemm rc=$? $(tail -1 gpurun_out/k_gemm.log)"
  timeout 300 $PT tests/test_layers_gpu.py > gpurun_out/l_all.log 2>&1; echo "l_all rc=$? $(tail -1 gpurun_out/l_all.log)"
  timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench1 rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 tools/dp_check.py > gpurun_out/dp_check_${NG:-2}.log 2>&1; echo "dp_check rc=$?"; tail -3 gpurun_out/dp_check_${NG:-2}.log
timeout 300 $TR --master-port 29502 bench.py --gpus ${NG:-2} --steps 20 --warmup 5 > gpurun_out/bench_n${NG:-2}.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n${NG:-2}.log | cut -c1-400
timeout 300 $TR --master-port 29503 tools/dp_timeline.py > gpurun_out/dp_timeline_n${NG:-2}.log 2>&1; echo "timeline rc=$?"; grep -v "Warning\|warn" gpurun_out/dp_timeline_n${NG:-2}.log | tail -62
