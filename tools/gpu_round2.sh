mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 600 tests/test_gpu_bf16.py -k "overfits" > gpurun_out/overfit.log 2>&1
tail -n 12 gpurun_out/overfit.log
