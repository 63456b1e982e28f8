#!/bin/bash
# ncu evidence for profiles/: (1) the launch list of a short bench run, (2) one --set full capture of
# the GEMM kernels of one training step.  Each ncu run follows the same command run plain (exit 0).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile"
$CMD > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/profile_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 22 -c 11 \
    -f -o gpurun_out/prof_gemm $CMD > gpurun_out/profile_ncu2.log 2>&1
echo "gemm capture rc=$?"
tail -n 3 gpurun_out/profile_ncu2.log
