#!/bin/bash
# Runs the GPU test groups in separate processes (a faulting kernel poisons its CUDA context, so
# groups are isolated), then smoke() and a short bench.  Everything is logged under gpurun_out/.
# Usage (on the GPU box, from the repo root): bash tools/gpu_ci.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 gpurun_out/$name.log | cut -c1-200)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run k_misc      600 $PT tests/test_kernels_gpu.py -k "not gemm and not variants"
run k_popc      300 $PT tests/test_kernels_gpu.py -k "gemm_popc"
run k_i8_simt   600 $PT tests/test_kernels_gpu.py -k "gemm_i8_exact and simt"
run k_f16_simt  600 $PT tests/test_kernels_gpu.py -k "gemm_f16 and simt"
run k_i8_tc     300 $PT tests/test_kernels_gpu.py -k "gemm_i8_exact and tcgen05"
run k_i8_tc2    300 $PT tests/test_kernels_gpu.py -k "unaligned_output_and_batch or rejects_bad or variants_agree"
run k_f16_tc    300 $PT tests/test_kernels_gpu.py -k "gemm_f16 and tcgen05"
run l_simt      900 $PT tests/test_layers_gpu.py -k "simt or batchnorm_layer_api or image_error"
run l_tc        900 $PT tests/test_layers_gpu.py -k "not simt and not config2 and not batchnorm_layer_api and not image_error"
run l_config2   600 $PT tests/test_layers_gpu.py -k "config2"
run smoke       300 python -c "import __graft_entry__ as g; g.smoke()"
run bench       600 python bench.py --steps 10 --warmup 3
cat gpurun_out/summary.txt
