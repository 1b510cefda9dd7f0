# Round 2: full GPU test suite (incl. batch-256 selfcheck, C driver), memcheck, reference arms, bench with extras
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/r2f_pytest.log; grep -E "selfcheck|batch-256 TF32" gpurun_out/r2f_pytest.log
bash tools/gpu_sanitize.sh memcheck > gpurun_out/r2f_sanitize.log 2>&1; cat gpurun_out/r02_sanitizer_memcheck.txt
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/r2f_ref_c2.json 2> gpurun_out/r2f_ref_c2.err; echo "ref c2 exit $?"
timeout 600 python bench.py --impl reference_cached --steps 6 --warmup 3 > gpurun_out/r2f_refcached_c2.json 2> gpurun_out/r2f_refcached_c2.err; echo "ref cached c2 exit $?"
timeout 600 python bench.py --impl reference --config c3 --steps 4 --warmup 3 > gpurun_out/r2f_ref_c3.json 2> gpurun_out/r2f_ref_c3.err; echo "ref c3 exit $?"
timeout 600 python bench.py --impl reference_cached --config c3 --steps 4 --warmup 3 > gpurun_out/r2f_refcached_c3.json 2> gpurun_out/r2f_refcached_c3.err; echo "ref cached c3 exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
for f in r2f_ref_c2 r2f_refcached_c2 r2f_ref_c3 r2f_refcached_c3; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d.get('impl'), d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('reference_gpu_error'))
"; done
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('value'), v.get('error'))
"
