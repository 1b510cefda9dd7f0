mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
for c in c2 c4; do
timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
tail -n 3 gpurun_out/pytest_gpu.log
