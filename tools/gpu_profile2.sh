#!/bin/bash
# Binary-GEMM variant comparison (timing + ncu counters) and the HBM kernels' full capture.
mkdir -p gpurun_out
python tools/bench_binary_gemm.py > gpurun_out/binary_gemm.json 2> gpurun_out/binary_gemm.err
echo "binary gemm rc=$?"; cat gpurun_out/binary_gemm.json
python tools/bench_binary_gemm.py --reps 2 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_(tc|popc)_kernel' -s 6 -c 2 -f \
    -o gpurun_out/prof_binary python tools/bench_binary_gemm.py --reps 2 > gpurun_out/prof_binary.log 2>&1
echo "ncu binary rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --profile"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bn_|colstats|split_f16|binarize' -s 46 -c 34 -f \
    -o gpurun_out/prof_hbm $CMD > gpurun_out/prof_hbm.log 2>&1
echo "ncu hbm rc=$?"
