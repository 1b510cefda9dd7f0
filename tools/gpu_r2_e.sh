mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --dtype f32 --variants "RESNET_B200_PRODUCERS=1;RESNET_B200_PRODUCERS=4" > gpurun_out/r2e_prod_f32.txt 2>&1; echo "prod f32 exit $?"
timeout 300 python tools/conv_bench.py --dtype bf16 --variants "RESNET_B200_PRODUCERS=1;RESNET_B200_PRODUCERS=4" > gpurun_out/r2e_prod_bf16.txt 2>&1; echo "prod bf16 exit $?"
tail -n 1 gpurun_out/r2e_prod_f32.txt gpurun_out/r2e_prod_bf16.txt
