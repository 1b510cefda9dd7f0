mkdir -p gpurun_out
timeout 120 ./tools/mma_rate2.bin > gpurun_out/r02_mma_rate2.txt 2>&1; echo "mma_rate2 exit $?"; cat gpurun_out/r02_mma_rate2.txt
timeout 300 python tools/conv_bench.py --dtype f32 > gpurun_out/r2h_conv_f32.txt 2>&1; echo "conv f32 exit $?"
timeout 300 python tools/conv_bench.py --dtype bf16 > gpurun_out/r2h_conv_bf16.txt 2>&1; echo "conv bf16 exit $?"
tail -n 1 gpurun_out/r2h_conv_f32.txt gpurun_out/r2h_conv_bf16.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r2h_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('value'), v.get('error'))
"
