mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --dtype f32 --only "1x1" --passes 0,1 --variants "RESNET_B200_PF_MIN_ROW=0;RESNET_B200_PF_MIN_ROW=1024" > gpurun_out/r2m_pf_f32.txt 2>&1; echo "f32 exit $?"
timeout 300 python tools/conv_bench.py --dtype bf16 --only "1x1" --passes 0,1 --variants "RESNET_B200_PF_MIN_ROW=0;RESNET_B200_PF_MIN_ROW=1024" > gpurun_out/r2m_pf_bf16.txt 2>&1; echo "bf16 exit $?"
cat gpurun_out/r2m_pf_f32.txt; tail -n 1 gpurun_out/r2m_pf_bf16.txt
