# Round 2, first GPU call: re-measure the MMA issue rate with warp-uniform issue loops, run the GPU tests, time every conv layer.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 120 ./tools/mma_rate.bin > gpurun_out/r02_mma_rate.txt 2>&1; echo "mma_rate exit $?"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r2a_pytest.log
timeout 300 python tools/conv_bench.py --dtype bf16 --variants "RESNET_B200_HALO=1" > gpurun_out/r2a_conv_bf16.txt 2>&1; echo "conv_bench bf16 exit $?"
timeout 300 python tools/conv_bench.py --dtype f32 --variants "RESNET_B200_HALO=1" > gpurun_out/r2a_conv_f32.txt 2>&1; echo "conv_bench f32 exit $?"
timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench c2 exit $?"
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; echo "bench c4 exit $?"
tail -n 12 gpurun_out/r02_mma_rate.txt
tail -n 3 gpurun_out/r2a_conv_bf16.txt gpurun_out/r2a_conv_f32.txt
cat gpurun_out/r2a_bench_c2.json gpurun_out/r2a_bench_c4.json | cut -c1-400
