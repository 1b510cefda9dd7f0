mkdir -p gpurun_out
for shp in 1,1,1024,256,14 1,1,256,1024,14 1,1,2048,512,7 1,1,512,2048,7 3,1,256,256,14 3,1,512,512,7 1,1,512,256,28; do
timeout 120 python tools/conv_bench.py --dtype f32 --shape $shp --passes 0,1 --desc --variants "RESNET_B200_DEBUG_SKIP=1;RESNET_B200_DEBUG_SKIP=3;RESNET_B200_STAGES=2" 2>&1 | grep -v "^#"
done > gpurun_out/r2j_feed_f32.txt
cat gpurun_out/r2j_feed_f32.txt
