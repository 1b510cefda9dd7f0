# ncu launch lists of one ResNet-50 step in both storage modes (profiles/r02_ncu_all_kernels_*) and of the bench command itself: run on
# a B200 box, then
#   python tools/ncu_summary.py gpurun_out/ncu_all_<dt>.csv --trace gpurun_out/trace_<dt>.log --last-step --json profiles/r02_traffic_<dt>.json
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for dt in tf32 bf16; do
RESNET_B200_TRACE=1 timeout 300 python tools/one_step.py --dtype $dt --steps 2 > gpurun_out/plain_$dt.log 2> gpurun_out/trace_$dt.log &&
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_all_$dt.csv python tools/one_step.py --dtype $dt --steps 2 > gpurun_out/ncu_$dt.log 2>&1
echo "ncu $dt exit $?"
done
# the bench command's own launch list (durations only): the kernel shares of the step the JSON line reports
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_steps2.json 2> gpurun_out/bench_steps2.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench exit $?"
ls -la gpurun_out/ncu_all_*.csv gpurun_out/launches_bench_steps2.csv
