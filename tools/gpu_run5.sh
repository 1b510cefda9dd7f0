set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16.py -q -x --timeout 120 -k "conv or wgrad or stem or halo or resident" > gpurun_out/pytest_iss2.log 2>&1
tail -n 3 gpurun_out/pytest_iss2.log
RESNET_B200_ISSUERS=4 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bf16.py -q -x --timeout 120 -k "conv or wgrad or stem or halo or resident" > gpurun_out/pytest_iss4.log 2>&1
tail -n 3 gpurun_out/pytest_iss4.log
timeout 300 python tools/probe_issuers.py > gpurun_out/issuers.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/issuers_ncu.csv -k regex:igemm python tools/probe_issuers.py > gpurun_out/issuers_ncu.log 2>&1
tail -n 3 gpurun_out/issuers.log
