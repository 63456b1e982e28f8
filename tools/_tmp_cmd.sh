mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 300 $PT tests/test_layers_gpu.py > gpurun_out/l_all.log 2>&1; echo "l_all rc=$? $(tail -1 gpurun_out/l_all.log)"; grep -E "Error|FAILED" gpurun_out/l_all.log | head -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('N1', d['ms_per_step'], d['ms_per_step_eager'], d['e2e']['ms_per_step'], d['e2e']['serial_train_step']['ms_per_step'], d['gpu_launches'])
PY
