mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for dt in bf16 tf32; do
timeout 300 python tools/one_step.py --dtype $dt > gpurun_out/plain_$dt.log 2>&1 &&
timeout 600 ncu --metrics $M --clock-control none -k regex:maxpool --csv --log-file gpurun_out/ncu_pool_$dt.csv python tools/one_step.py --dtype $dt > gpurun_out/ncu_pool.log 2>&1
done
