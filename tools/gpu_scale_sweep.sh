# weak-scaling sweep 1 / 2 / 4 / 8 GPUs on ONE 8-GPU box (same silicon, same thermals): gpurun --gpus 8 -- 'bash tools/gpu_scale_sweep.sh'
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err
  fi
  echo "N=$N exit $?"
done
python - <<'PY'
import json
base = {}
for N in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open('gpurun_out/sweep_n%d.json' % N) if l.startswith('{')][-1])
    except Exception as e:
        print(N, 'FAILED', e); continue
    row = {'c2': (d['value'], d['ms_per_step'])}
    for k, v in d.get('extra', {}).items():
        if 'value' in v: row[k] = (v['value'], v['ms_per_step'])
    if N == 1: base = row
    print('N=%d ' % N + '  '.join('%s %.0f img/s %.2f ms (x%.2f)' % (k, v[0], v[1], v[0] / base[k][0] if k in base else 0) for k, v in row.items()), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
PY
