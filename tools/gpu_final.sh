set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for c in c3 c4 c5; do
timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
tail -n 2 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
