mkdir -p gpurun_out
free -g | head -2 > gpurun_out/parity_batch32.txt
timeout 1200 python tools/parity_batch.py 32 >> gpurun_out/parity_batch32.txt 2>&1
tail -n 20 gpurun_out/parity_batch32.txt
