mkdir -p gpurun_out
for cfg in c2 c4; do
for iss in 1 2 4; do
RESNET_B200_ISSUERS=$iss timeout 200 python bench.py --config $cfg --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/iss_${cfg}_${iss}.json 2> gpurun_out/iss_${cfg}_${iss}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/iss_${cfg}_${iss}.json").read().strip().splitlines()[-1])
print("$cfg issuers=$iss halo=0", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"], d["roofline"]["unit"])
PY
done
for iss in 2 4; do
RESNET_B200_HALO=1 RESNET_B200_ISSUERS=$iss timeout 200 python bench.py --config $cfg --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/iss_${cfg}_${iss}_halo.json 2> gpurun_out/iss_${cfg}_${iss}_halo.err
python - <<PY
import json
d=json.loads(open("gpurun_out/iss_${cfg}_${iss}_halo.json").read().strip().splitlines()[-1])
print("$cfg issuers=$iss halo=1", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"], d["roofline"]["unit"])
PY
done
done
