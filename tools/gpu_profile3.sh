#!/bin/bash
# ncu evidence for profiles/: launch list of one eager bench run, then --set full captures of the GEMM
# launches and of the HBM-bound kernels of one training step.  Each capture follows a plain run of the
# same command that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/profile_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel' -s 44 -c 11 -f \
    -o /tmp/prof_gemm $CMD > gpurun_out/prof_gemm.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'bn_|colstats|split_f16|binarize|absmax|softmax|sgd' -s 132 -c 33 -f \
    -o /tmp/prof_hbm $CMD > gpurun_out/prof_hbm.log 2>&1
echo "ncu hbm rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed"
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > gpurun_out/prof_hbm_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/*raw.csv
# the .ncu-rep files stay on the box: gpurun_out/ is limited to 64 MiB
