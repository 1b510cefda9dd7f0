set -x
mkdir -p gpurun_out
timeout 120 ./tools/mma_rate.bin > gpurun_out/mma_rate.txt 2>&1
bash tools/gpu_final.sh
bash tools/gpu_profile.sh
