mkdir -p gpurun_out
timeout 600 python tools/probe_resident.py > gpurun_out/probe2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:igemm --csv --log-file gpurun_out/probe2_ncu.csv python tools/probe_resident.py > gpurun_out/probe2_ncu.log 2>&1
