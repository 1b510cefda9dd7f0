mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --dtype f32 --variants "RESNET_B200_L2_PROMO=256;RESNET_B200_L2_PROMO=64" > gpurun_out/r2n_promo_f32.txt 2>&1; echo "f32 exit $?"
timeout 300 python tools/conv_bench.py --dtype bf16 --variants "RESNET_B200_L2_PROMO=256;RESNET_B200_L2_PROMO=64" > gpurun_out/r2n_promo_bf16.txt 2>&1; echo "bf16 exit $?"
grep -E "1x1|#" gpurun_out/r2n_promo_f32.txt; tail -n 1 gpurun_out/r2n_promo_bf16.txt
