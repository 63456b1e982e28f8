#!/bin/bash
# One entry point for everything that runs on the GPU box (replaces the one-off tools/gpu_*.sh of round 1).
#   gpurun --timeout 900 -- 'bash tools/gpu.sh <task> [<task> ...]'
# Every task runs in its own process under its own `timeout` (a trapped kernel poisons its CUDA context,
# a hung one must not take the box with it) and logs to gpurun_out/<name>.log; one summary line per step
# goes to stdout and gpurun_out/summary.txt.
#
# tasks
#   ci            the GPU test groups (separate processes), smoke(), a short bench
#   tests         `pytest -m gpu` in one process, as the driver runs it
#   parity        the BASELINE-config parity tests (tests/test_config_parity_gpu.py)
#   pair          CTA-pair GEMM bring-up: opt-in tests, then the bench with the pair form on / off
#   mc            cluster-multicast GEMM bring-up: opt-in tests, then benches with the form on / off
#   probe         tools/probe_mxf4.cu: +-1 products on the FP4 tensor pipe (exactness + MMA rate)
#   bench         bench.py at N=1 (BENCH_ARGS overrides "--steps 20 --warmup 5")
#   bench-long    bench.py --steps 300 (sustained clocks)
#   reference     bench.py --impl reference
#   launches      ncu launch list of one eager training step -> gpurun_out/launches.csv
#   prof-gemm     ncu --set full of the GEMM launches of one training step -> gpurun_out/prof_gemm_raw.csv
#   prof-hbm      ncu --set full of the HBM-bound kernels of one step      -> gpurun_out/prof_hbm_raw.csv
#   prof-sweep    ncu --set full of the fault-sweep kernels                 -> gpurun_out/prof_sweep_raw.csv
#   walk          tile-walk A/B (BNN_GEMM_GN / BNN_GEMM_GM) on the headline step
#   sweep         tools/bench_sweep.py (configs 4 / 5)
#   sanitizer     compute-sanitizer memcheck / racecheck / synccheck on small-shape kernel tests
#   dp N          tools/dp_check.py + bench.py on N GPUs (needs gpurun --gpus N)
mkdir -p gpurun_out
: "${BENCH_ARGS:=--steps 20 --warmup 5}"
PT="python -m pytest -q -m gpu -p no:cacheprovider --tb=short"
PROFILE_CMD="python bench.py --steps 2 --warmup 3 --profile"

run() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc  $(tail -n 1 "gpurun_out/$name.log" | cut -c1-220)" | tee -a gpurun_out/summary.txt
  return $rc
}

TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

raw_csv() {  # report base name
  ncu -i "/tmp/$1.ncu-rep" --page raw --csv > "gpurun_out/$1_raw.csv" 2>/dev/null
  ls -la "/tmp/$1.ncu-rep" "gpurun_out/$1_raw.csv"
}

nvidia-smi --query-gpu=name,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
while [ $# -gt 0 ]; do
  task=$1; shift
  echo "=== $task" | tee -a gpurun_out/summary.txt
  case "$task" in
    ci)
      run k_misc     600 $PT tests/test_kernels_gpu.py -k "not gemm and not variants"
      run k_gemm     900 $PT tests/test_kernels_gpu.py -k "gemm or variants"
      run l_layers   900 $PT tests/test_layers_gpu.py -k "not config2"
      run l_config2  600 $PT tests/test_layers_gpu.py -k "config2"
      run u8         600 $PT tests/test_u8_input_gpu.py
      run parity     900 $PT -s tests/test_config_parity_gpu.py
      run smoke      300 python -c "import __graft_entry__ as g; g.smoke()"
      run bench      600 python bench.py --steps 10 --warmup 3
      ;;
    tests)   run tests 1800 $PT tests ;;
    parity)  run parity 900 $PT -s tests/test_config_parity_gpu.py ;;
    pair)
      export BNN_TEST_CTA2=1
      PP="$PT -x tests/test_gemm_pair_gpu.py"
      run pair_i8_small  60 $PP -k "int8 and 256-256-128" &&
      run pair_i8_all   120 $PP -k "int8" &&
      run pair_f16      120 $PP -k "fp16" &&
      run pair_step     120 $PP -k "training_step" && {
        BNN_GEMM_CTA2=0 run pair_bench_off 240 python bench.py $BENCH_ARGS
        BNN_GEMM_CTA2=1 run pair_bench_on  240 python bench.py $BENCH_ARGS
      }
      unset BNN_TEST_CTA2
      ;;
    mc)
      export BNN_TEST_MC=1
      PM="$PT -x tests/test_gemm_multicast_gpu.py"
      run mc_i8_small  60 $PM -k "int8 and 256-256-128" &&
      run mc_i8_all   120 $PM -k "int8" &&
      run mc_f16      120 $PM -k "fp16" &&
      run mc_fp4      120 $PM -k "fp4" &&
      run mc_step     120 $PM -k "training_step" && {
        BNN_GEMM_MC=0 run mc_bench_off 240 python bench.py $BENCH_ARGS --no-extras --no-cpu
        BNN_GEMM_MC=1 run mc_bench_on  240 python bench.py $BENCH_ARGS --no-extras --no-cpu
        BNN_GEMM_MC=0 FORMS_SINGLE_ONLY=1 FORMS_CASES=f4_sign_f4,i8_store_s16,i8_sign_i8,f16_1pass_f32,f16_3pass_f32,k784_i8_sign_f4 run mc_forms_off 200 python tools/bench_gemm_forms.py
        BNN_GEMM_MC=1 FORMS_SINGLE_ONLY=1 FORMS_CASES=f4_sign_f4,i8_store_s16,i8_sign_i8,f16_1pass_f32,f16_3pass_f32,k784_i8_sign_f4 run mc_forms_on 200 python tools/bench_gemm_forms.py
      }
      unset BNN_TEST_MC
      ;;
    probe)
      run probe_build 120 nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/probe_mxf4.cu -o /tmp/probe_mxf4 &&
      run probe_mxf4   60 /tmp/probe_mxf4
      ;;
    bench)       run bench 600 python bench.py $BENCH_ARGS ;;
    bench-long)  run bench_long 600 python bench.py --steps 300 --warmup 20 ;;
    reference)   run bench_reference 600 python bench.py --impl reference $BENCH_ARGS ;;
    launches)
      run profile_plain 300 $PROFILE_CMD &&
      run launches_ncu 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
          --log-file gpurun_out/launches.csv $PROFILE_CMD
      ;;
    prof-gemm)
      # the eager profile run does 3 warm-up + 2 timed steps of 11 GEMM launches: skip the first 4 steps
      run profile_plain 300 $PROFILE_CMD &&
      run prof_gemm 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel' \
          -s 44 -c 11 -f -o /tmp/prof_gemm $PROFILE_CMD && raw_csv prof_gemm
      ;;
    prof-hbm)
      run profile_plain 300 $PROFILE_CMD &&
      run prof_hbm 900 ncu --set full --clock-control none --import-source on \
          -k regex:'bn_|colstats|split_f16|binarize|absmax|softmax|sgd|gather' -s 128 -c 32 -f \
          -o /tmp/prof_hbm $PROFILE_CMD && raw_csv prof_hbm
      ;;
    prof-sweep)
      # one launch group of 8 trials per sweep (plus the warm-up groups): flips, expansion, thresholds, GEMMs
      SW="python tools/bench_sweep.py --trials 8 --image-trials 8 --cpu-trials 0 --sections weights,image"
      run sweep_plain 300 $SW &&
      run prof_sweep 900 ncu --set full --clock-control none --import-source on \
          -k regex:'flip_|thresholds|bits_to_f4|gemm_tc|bn_apply|count_correct' -s 26 -c 40 -f -o /tmp/prof_sweep $SW && raw_csv prof_sweep
      ;;
    walk)
      # tile-walk A/B on the headline step: BNN_GEMM_GN = column tiles per super-column (16 = one super-column,
      # the round-1 walk), BNN_GEMM_GM = row tiles per group
      for cfg in "0 0" "16 8" "8 8" "4 8" "8 16" "8 4" "4 16"; do
        set -- $cfg "$@"; gn=$1; gm=$2; shift 2
        BNN_GEMM_GN=$gn BNN_GEMM_GM=$gm run "walk_gn${gn}_gm${gm}" 240 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu
      done
      ;;
    sweep)  run sweep 900 python tools/bench_sweep.py --sections config1,weights,image,inference ;;
    sanitizer)
      ST="$PT tests/test_kernels_gpu.py -k 'not full_size and not 8192 and not 4096 and not simt'"
      for tool in memcheck racecheck synccheck; do
        run "sanitizer_$tool" 1500 bash -c "compute-sanitizer --tool $tool --error-exitcode 9 $ST"
      done
      ;;
    dp)
      n=$1; shift
      run "dp_check_$n" 600 $TORCHRUN --nproc-per-node "$n" --master-port 29511 tools/dp_check.py
      run "bench_n$n"   600 $TORCHRUN --nproc-per-node "$n" --master-port 29512 bench.py --gpus "$n" $BENCH_ARGS
      ;;
    *) echo "unknown task $task"; exit 2 ;;
  esac
done
cat gpurun_out/summary.txt
