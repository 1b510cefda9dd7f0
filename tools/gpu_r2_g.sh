mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r2g_pytest.log
for v in 1 0 1 0; do
for c in c2 c4; do
RESNET_B200_ASYNC_WGRAD=$v timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_${c}_async$v.json 2> gpurun_out/r2g_${c}_async$v.err
python -c "
import json
d=json.loads(open('gpurun_out/r2g_${c}_async$v.json').read().strip().splitlines()[-1])
print('$c async=$v', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['last_step'])
"
done
done
