# ncu launch lists of one ResNet-50 step in both storage modes (profiles/r01_ncu_all_kernels_*): run on a B200 box, then
#   python tools/ncu_summary.py gpurun_out/ncu_all_<dt>.csv --trace gpurun_out/trace_<dt>.log --last-step --json profiles/r01_traffic_<dt>.json
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for dt in bf16 tf32; do
RESNET_B200_TRACE=1 timeout 300 python tools/one_step.py --dtype $dt > gpurun_out/plain_$dt.log 2> gpurun_out/trace_$dt.log &&
timeout 1200 ncu --metrics $M --clock-control none -s 300 --csv --log-file gpurun_out/ncu_all_$dt.csv python tools/one_step.py --dtype $dt > gpurun_out/ncu_$dt.log 2>&1
done
