set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 120 ./tools/mma_rate.bin > gpurun_out/mma_rate.txt 2>&1
echo "rc=$?" >> gpurun_out/mma_rate.txt
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16.py -q -x --timeout 300 -k "conv or wgrad or stem or halo or resident" > gpurun_out/pytest_run2.log 2>&1
tail -n 5 gpurun_out/pytest_run2.log
cat gpurun_out/mma_rate.txt
