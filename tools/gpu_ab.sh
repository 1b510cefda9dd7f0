mkdir -p gpurun_out
for c in c2 c4; do
for v in "1 1" "0 1" "1 0" "0 0" "1 1"; do
set -- $v
RESNET_B200_FUSED_BN_FWD=$1 RESNET_B200_FUSED_BN_BWD=$2 timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab.json'))
print('$c fwd=$1 bwd=$2', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['gpu_launches'], [round(r['ms_per_step'],2) for r in d['roofline_all']])
PY
done
done
