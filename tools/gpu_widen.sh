#!/bin/bash
# GPU check of the SURVEY 8(f) additions (uint8-input layer 0, IEEE-754 weight flips), one process per
# group so that a faulting kernel cannot mask the other groups.  Logs under gpurun_out/.
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --tb=short tests/test_u8_input_gpu.py"
run() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-160)"; }
run w_u8gemm_tc 45 $PT -k "gemm_u8_s8_exact and tcgen05"
run w_misc      45 $PT -k "(gemm_u8_s8_exact and simt) or rejects or byproducts or f32_bit or float_branch or biased"
run w_native    45 $PT -k "native or full_size"
run w_sweep     45 python tools/bench_sweep.py --sections image --image-trials 16 --cpu-trials 0
