#!/bin/bash
# After the flip-kernel rework: the driver's GPU test command, smoke(), and the image-sweep bench again.
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
run all2       90 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short
run smoke2     40 python -c "import __graft_entry__ as g; g.smoke()"
run sweep_img2 60 python tools/bench_sweep.py --sections image --image-trials 64 --cpu-trials 0
