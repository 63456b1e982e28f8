mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 300 $PT tests/test_kernels_gpu.py > gpurun_out/k_all.log 2>&1; echo "k_all rc=$? $(tail -1 gpurun_out/k_all.log)"; grep -E "^E  |FAILED" gpurun_out/k_all.log | head -12
timeout 300 $PT tests/test_layers_gpu.py > gpurun_out/l_all.log 2>&1; echo "l_all rc=$? $(tail -1 gpurun_out/l_all.log)"; grep -E "^E  |FAILED" gpurun_out/l_all.log | head -12
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log | cut -c1-200; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('N1', d['ms_per_step'], d['ms_per_step_eager'], d['e2e']['ms_per_step'], d['gpu_launches'], d['binary_gemm']['avg_launch_ms'], d['binary_gemm']['launches'], d['roofline']['avg_launch_ms'], d['roofline']['launches'])
PY
BNN_DW_INT8=0 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_noq.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_noq.log').read().strip().splitlines()[-1])
print('N1 fp16-dW', d['ms_per_step'], d['e2e']['ms_per_step'])
PY
