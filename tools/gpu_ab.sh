# same-box A/B of one environment switch: bash tools/gpu_ab.sh VAR "v1 v2" "c2 c4 c5" [reps]   (alternating runs, bench.py without extras)
VAR=$1; VALS=${2:-"0 1"}; CFGS=${3:-"c2 c4"}; REPS=${4:-2}
mkdir -p gpurun_out
for rep in $(seq 1 $REPS); do for cfg in $CFGS; do for v in $VALS; do
  env $VAR=$v timeout 600 python bench.py --config $cfg --no-extra --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/ab_${VAR}_${v}_${cfg}_$rep.json 2> gpurun_out/ab_${VAR}_${v}_${cfg}_$rep.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/ab_${VAR}_${v}_${cfg}_$rep.json').read().strip().splitlines()[-1])
    print('$cfg $VAR=$v rep $rep: %.1f img/s %.2f ms e2e %.1f clk %s launches %s %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'].get('sm_mhz'), d.get('gpu_launches'), d.get('last_step')))
except Exception as ex:
    print('$cfg $VAR=$v rep $rep: FAILED', ex)
P
done; done; done
