# end-of-round validation on a B200 box: smoke, GPU tests, both bench arms the driver runs (timed), extras
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/final_smoke.log
( time timeout 900 python -m pytest tests -q -m gpu --timeout 300 ) > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/final_pytest.log
( time timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err ) 2>&1 | tail -n 3; echo "ref exit $?"
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | tail -n 3; echo "bench exit $?"
python -c "
import json
r=json.loads(open('gpurun_out/final_bench_ref.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('ms_per_step'), r.get('e2e',{}).get('value'), r.get('config')==json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]).get('config'))
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
for r in d['roofline_all']: print('   ', r['kernel'][:48], round(r['ms_per_step'],2), round(r['achieved'],1), r['unit'], round(r['frac'],3))
for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('value'), v.get('error'))
print(d.get('cpu_baseline'))
"
