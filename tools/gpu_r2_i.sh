mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --dtype f32 --only "1x1" --passes 0,1 --variants "RESNET_B200_NSTAGING=3;RESNET_B200_NSTAGING=4;RESNET_B200_EPI_GROUPS=1,RESNET_B200_NSTAGING=4" > gpurun_out/r2i_1x1_f32.txt 2>&1; echo "f32 exit $?"
timeout 300 python tools/conv_bench.py --dtype bf16 --only "1x1" --passes 0,1 --variants "RESNET_B200_NSTAGING=3;RESNET_B200_NSTAGING=4;RESNET_B200_EPI_GROUPS=1,RESNET_B200_NSTAGING=4" > gpurun_out/r2i_1x1_bf16.txt 2>&1; echo "bf16 exit $?"
cat gpurun_out/r2i_1x1_f32.txt gpurun_out/r2i_1x1_bf16.txt
