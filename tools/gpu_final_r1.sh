#!/bin/bash
# Round-end style check on one B200: the driver's GPU test command, smoke(), the default bench line and
# the image-sweep bench (with its per-call timeline and the CPU port of one reference trial).
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
run all       120 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short
run smoke      60 python -c "import __graft_entry__ as g; g.smoke()"
run bench_n1  120 python bench.py
run sweep_img  90 python tools/bench_sweep.py --sections image --image-trials 64 --cpu-trials 1
