set -x
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q --timeout 300"
$T tests -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -n 3 gpurun_out/pytest_gpu.log
