set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
T="timeout 600 python -m pytest -q --timeout 240"
$T tests/test_gpu_bf16.py -k "conv_bf16_vs_oracle" > gpurun_out/bf16_conv.log 2>&1
$T tests/test_gpu_bf16.py -k "stem or linearity" > gpurun_out/bf16_stem.log 2>&1
$T tests/test_gpu_bf16.py -k "batchnorm or pools" > gpurun_out/bf16_bn.log 2>&1
$T tests/test_gpu_bf16.py -k "step or forward_only" > gpurun_out/bf16_net.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_bf16.py > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -3 gpurun_out/bf16_*.log gpurun_out/pytest_gpu.log
