#!/bin/bash
# Last check of round 1: the driver's GPU test command and the weight-sweep bench (fp32 and uint8 inputs).
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; }
run all3       60 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short
run sweep_w3   30 python tools/bench_sweep.py --sections weights --trials 32 --cpu-trials 0
