mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 600 $PT tests/test_layers_gpu.py > gpurun_out/l_all.log 2>&1; echo "l_all rc=$? $(tail -1 gpurun_out/l_all.log)"; grep -E "Error|assert|FAILED" gpurun_out/l_all.log | head -20
