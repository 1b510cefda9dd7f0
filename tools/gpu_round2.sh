mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
RESNET_B200_TRACE=1 timeout 300 python tools/one_step.py --dtype bf16 > gpurun_out/plain_bf16.log 2> gpurun_out/trace_apply.log &&
timeout 600 ncu --metrics $M --clock-control none -k regex:bn_apply --csv --log-file gpurun_out/ncu_apply.csv python tools/one_step.py --dtype bf16 > gpurun_out/ncu_apply.log 2>&1
