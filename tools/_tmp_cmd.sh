mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 tools/dp_check.py > gpurun_out/dp_check_2.log 2>&1; echo "dp_check rc=$?"; grep -v "Warning\|warn" gpurun_out/dp_check_2.log | tail -7
