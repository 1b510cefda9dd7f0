set -x
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q --timeout 300"
$T tests/test_gpu_dump.py tests/test_gpu_loader.py > gpurun_out/dump_tests.log 2>&1
for c in c2 c4 c3 c5; do
timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
tail -n 3 gpurun_out/dump_tests.log
