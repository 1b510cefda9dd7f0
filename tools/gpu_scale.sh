set -x
mkdir -p gpurun_out
N=$1
for c in c2 c4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c --steps 20 --warmup 5 > gpurun_out/scale_${c}_n$N.json 2> gpurun_out/scale_${c}_n$N.err
done
tail -n 2 gpurun_out/scale_*_n$N.json | cut -c1-400
