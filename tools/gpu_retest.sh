#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
  echo "rc=$rc  $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run k_misc      200 $PT tests/test_kernels_gpu.py -k "not gemm and not variants and not precision"
run k_gemm      200 $PT tests/test_kernels_gpu.py -k "gemm or variants or precision"
run l_all       300 $PT tests/test_layers_gpu.py
run smoke       120 python -c "import __graft_entry__ as g; g.smoke()"
run bench       300 python bench.py --steps 20 --warmup 5
CMD="python bench.py --steps 2 --warmup 3 --profile"
$CMD > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/profile_ncu1.log 2>&1
echo "launch list rc=$?"
