#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
  echo "rc=$rc  $(tail -n 1 gpurun_out/$name.log | cut -c1-200)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run k_f16_simt  600 $PT tests/test_kernels_gpu.py -k "gemm_f16 and simt"
run k_f16_tc    300 $PT -s tests/test_kernels_gpu.py -k "(gemm_f16 and tcgen05) or precision_at_full"
run l_tc        900 $PT tests/test_layers_gpu.py -k "not simt and not config2 and not batchnorm_layer_api and not image_error"
run smoke       300 python -c "import __graft_entry__ as g; g.smoke()"
