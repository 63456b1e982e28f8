#!/bin/bash
# Final single-GPU set: all GPU tests, smoke, bench, sweeps, launch list, and a --set full capture of
# the two kernels introduced last (int8 dW GEMM, quantizing BN-backward pass).
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider"
timeout 400 $PT tests/test_kernels_gpu.py > gpurun_out/k_all.log 2>&1; echo "k_all rc=$? $(tail -1 gpurun_out/k_all.log)"
timeout 400 $PT tests/test_layers_gpu.py > gpurun_out/l_all.log 2>&1; echo "l_all rc=$? $(tail -1 gpurun_out/l_all.log)"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-250
timeout 400 python tools/bench_sweep.py --trials 64 --cpu-trials 1 > gpurun_out/sweep_n1.log 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/sweep_n1.log | cut -c1-900
CMD="python bench.py --steps 2 --warmup 3 --profile"
$CMD > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/profile_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel<1, 256, 3|bn_bwd_apply_q' -s 12 -c 4 -f \
    -o /tmp/prof_new $CMD > gpurun_out/prof_new.log 2>&1
echo "ncu new rc=$?"
ncu -i /tmp/prof_new.ncu-rep --page raw --csv > gpurun_out/prof_new_raw.csv 2>/dev/null
ls -la gpurun_out/prof_new_raw.csv
