mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --dtype f32 --passes 0,1 --variants "RESNET_B200_L2_PREFETCH=0;RESNET_B200_L2_PREFETCH=24" > gpurun_out/r2k_pf_f32.txt 2>&1; echo "f32 exit $?"
timeout 300 python tools/conv_bench.py --dtype bf16 --passes 0,1 --variants "RESNET_B200_L2_PREFETCH=0;RESNET_B200_L2_PREFETCH=24" > gpurun_out/r2k_pf_bf16.txt 2>&1; echo "bf16 exit $?"
cat gpurun_out/r2k_pf_f32.txt; tail -n 1 gpurun_out/r2k_pf_bf16.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r2k_pytest.log
