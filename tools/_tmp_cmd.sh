mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 300 $PT tests/test_layers_gpu.py > gpurun_out/l_all.log 2>&1; echo "l_all rc=$? $(tail -1 gpurun_out/l_all.log)"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log | cut -c1-600; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('N1', d['ms_per_step'], d['ms_per_step_eager'], d['e2e']['ms_per_step'], d['e2e']['serial_train_step']['ms_per_step'], d['host_enqueue_ms_per_step'], d['host_enqueue_ms_per_step_eager'])
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 tools/dp_check.py > gpurun_out/dp_check_2.log 2>&1; echo "dp_check rc=$?"; tail -3 gpurun_out/dp_check_2.log
timeout 300 $TR --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?";  tail -3 gpurun_out/bench_n2.log | cut -c1-300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.log').read().strip().splitlines()[-1])
print('N2', d['ms_per_step'], d['ms_per_step_eager'], d['e2e']['ms_per_step'], d['e2e']['serial_train_step']['ms_per_step'], d['host_enqueue_ms_per_step'], d['host_enqueue_ms_per_step_eager'])
PY
