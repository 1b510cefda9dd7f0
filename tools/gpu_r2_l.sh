mkdir -p gpurun_out
for shp in 1,1,1024,256,14 1,1,2048,512,7 3,1,256,256,14 3,2,512,1024,28 1,1,512,256,28; do
timeout 120 python tools/conv_bench.py --dtype f32 --shape $shp --passes 0 --variants "RESNET_B200_DEBUG_SKIP=3;RESNET_B200_DEBUG_SKIP=7" 2>&1 | grep -v "^#"
done > gpurun_out/r2l_feed_ab_f32.txt
cat gpurun_out/r2l_feed_ab_f32.txt
