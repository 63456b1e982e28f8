mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 300 $PT tests/test_kernels_gpu.py > gpurun_out/k_all.log 2>&1; echo "k_all rc=$? $(tail -1 gpurun_out/k_all.log)"
timeout 300 $PT tests/test_layers_gpu.py > gpurun_out/l_all.log 2>&1; echo "l_all rc=$? $(tail -1 gpurun_out/l_all.log)"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench.log | cut -c1-300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('N1', d['ms_per_step'], d['ms_per_step_eager'], d['e2e']['ms_per_step'], d['e2e']['serial_train_step']['ms_per_step'], d['gpu_launches'])
PY
BNN_OVERLAP_DW=0 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_noov.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_noov.log').read().strip().splitlines()[-1])
print('N1 no-overlap', d['ms_per_step'], d['ms_per_step_eager'], d['e2e']['ms_per_step'])
PY
