# weak-scaling line (default bench = config 2 + the c4 / c5 extras) and the data-parallel numerics check on N GPUs of one box:
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_scale.sh N'
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py > gpurun_out/dp_check_n$N.log 2>&1; echo "dp_check exit $?"; grep -E "dp_check ok|rank 0" gpurun_out/dp_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "bench exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_n$N.json') if l.startswith('{')][-1])
print('c2', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1))
for k,v in d.get('extra',{}).items(): print(k, round(v.get('value',0),1), round(v.get('ms_per_step',0),2), v.get('error'))
"
