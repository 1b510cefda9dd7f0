set -x
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q --timeout 300"
$T tests -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
for c in c2 c4; do
timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for dt in bf16; do
RESNET_B200_TRACE=1 timeout 300 python tools/one_step.py --dtype $dt > gpurun_out/plain_$dt.log 2> gpurun_out/trace_$dt.log &&
timeout 1200 ncu --metrics $M --clock-control none -s 300 --csv --log-file gpurun_out/ncu_all_$dt.csv python tools/one_step.py --dtype $dt > gpurun_out/ncu_$dt.log 2>&1
done
