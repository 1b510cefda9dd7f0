# ncu --set full captures of representative convolution launches (one launch each, after a plain run of the same command)
mkdir -p gpurun_out
prof() {  # name dtype only pass
  local cmd="python tools/conv_bench.py --dtype $2 --shape $3 --passes $4 --iters 2 --warmup 1"
  $cmd > gpurun_out/ncu_layer_plain_$1.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_ -s 1 -c 1 -f -o gpurun_out/ncu_layer_$1 $cmd > gpurun_out/ncu_layer_ncu_$1.log 2>&1
  echo "$1 exit $?"
}
prof f1x1_256_1024 f32 1,1,256,1024,14 0
prof f3x3_64 f32 3,1,64,64,56 0
prof f3x3_256 f32 3,1,256,256,14 0
prof w1x1_1024_256 f32 1,1,1024,256,14 3
prof d3x3s2_256_512 f32 3,2,256,512,56 1
prof f1x1_64_256_bf16 bf16 1,1,64,256,56 0
ls -la gpurun_out/*.ncu-rep
