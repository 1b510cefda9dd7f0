for dt in f32 bf16; do
python tools/conv_bench.py --dtype $dt --passes 0,1 --only 1x1 --variants "RESNET_B200_MAX_BN=128,RESNET_B200_TWO_CTA=128,RESNET_B200_TWO_CTA_MINK=1;RESNET_B200_TWO_CTA=64,RESNET_B200_TWO_CTA_MINK=1" 2>&1 | cut -c1-150
done
