#!/bin/bash
# Bring-up of the CTA-pair GEMM on one B200: every group in its own process (a trapped kernel
# poisons its CUDA context), smallest shapes first, then the bench with and without the pair form, then an
# ncu capture of the pair kernel.  Logs under gpurun_out/pair_*.log.
#   gpurun --timeout 600 -- 'bash tools/gpu_pair_bringup.sh'
mkdir -p gpurun_out
export BNN_TEST_CTA2=1
PT="python -m pytest -q -m gpu -p no:cacheprovider --tb=short -x tests/test_gemm_pair_gpu.py"
run() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "$name rc=$rc $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; return $rc; }
run pair_i8_small   60 $PT -k "int8 and 256-256-128" || exit 1
run pair_i8_all    120 $PT -k "int8"                  || exit 1
run pair_f16       120 $PT -k "fp16"                  || exit 1
run pair_step      120 $PT -k "training_step"         || exit 1
BNN_GEMM_CTA2=0 run pair_bench_off 180 python bench.py --steps 20 --warmup 5
BNN_GEMM_CTA2=1 run pair_bench_on  180 python bench.py --steps 20 --warmup 5
BNN_GEMM_CTA2=1 BNN_STEP_GRAPH=0 run pair_ncu 300 ncu --set full --clock-control none --import-source on \
    -k regex:gemm_tc_kernel -c 12 -o gpurun_out/pair_gemm python bench.py --steps 1 --warmup 1
