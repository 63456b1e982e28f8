#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
  echo "rc=$rc  $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run k_i8_tc     120 $PT tests/test_kernels_gpu.py -k "gemm_i8_exact and tcgen05"
run k_i8_tc2    120 $PT tests/test_kernels_gpu.py -k "unaligned_output_and_batch or rejects_bad or variants_agree"
run k_f16_tc    120 $PT -s tests/test_kernels_gpu.py -k "(gemm_f16 and tcgen05) or precision_at_full"
BNN_GEMM_CHUNK=16 run k_prec16 90 $PT -s tests/test_kernels_gpu.py -k "precision_at_full"
BNN_GEMM_CHUNK=4 run k_prec4 90 $PT -s tests/test_kernels_gpu.py -k "precision_at_full"
run l_tc        240 $PT tests/test_layers_gpu.py -k "not simt and not batchnorm_layer_api and not image_error"
run smoke       120 python -c "import __graft_entry__ as g; g.smoke()"
run bench       300 python bench.py --steps 10 --warmup 3
BNN_GEMM_CHUNK=16 run bench16 120 python bench.py --steps 10 --warmup 3 --profile
BNN_GEMM_CHUNK=4 run bench4 120 python bench.py --steps 10 --warmup 3 --profile
BNN_GEMM_CHUNK=4096 run benchinf 120 python bench.py --steps 10 --warmup 3 --profile
grep -h "split-GEMM" gpurun_out/k_f16_tc.log gpurun_out/k_prec16.log gpurun_out/k_prec4.log
