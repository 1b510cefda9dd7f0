set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 python tools/probe_halo.py check > gpurun_out/halo_check.log 2>&1
echo "check rc=$?" >> gpurun_out/halo_check.log
timeout 300 python tools/probe_halo.py time > gpurun_out/halo_time.log 2>&1 &&
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/halo_time_ncu.csv -k regex:igemm python tools/probe_halo.py time > gpurun_out/halo_time_ncu.log 2>&1
tail -n 30 gpurun_out/halo_check.log
