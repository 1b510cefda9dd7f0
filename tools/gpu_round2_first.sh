# First GPU minutes of round 2: validate the slot-ownership multi-issuer mode and measure it.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_first.sh 2>&1 | tee gpurun_out/round2_first.txt'
# Every step runs under its own short timeout: the first version of this mode hung in the batch-256 step.
mkdir -p gpurun_out
echo "== unit tests with 1 / 2 / 4 issuers"
RESNET_B200_TEST_ISSUERS=1 timeout 200 python -m pytest tests/test_gpu_bf16.py -q --timeout 60 -k "issuer_counts" 2>&1 | tail -n 3
step() { echo "== one step, batch 256, $*"; env "$@" timeout 60 python tools/one_step.py --batch 256 --dtype bf16 --steps 2 2>&1 | grep -v "^\[" | tail -n 1; }
step RESNET_B200_ISSUERS_K=2
step RESNET_B200_ISSUERS_W=2
step RESNET_B200_ISSUERS=2
step RESNET_B200_ISSUERS=4
step RESNET_B200_ISSUERS=4 RESNET_B200_HALO=1
for cfg in c2 c4; do
for v in "RESNET_B200_ISSUERS=1" "RESNET_B200_ISSUERS=2" "RESNET_B200_ISSUERS=4" "RESNET_B200_ISSUERS_K=4 RESNET_B200_ISSUERS_W=1" "RESNET_B200_ISSUERS=4 RESNET_B200_HALO=1"; do
env $v timeout 120 python bench.py --config $cfg --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/r2_${cfg}.json 2> gpurun_out/r2_${cfg}.err
python - "$cfg" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], "%.1f img/s  %.2f ms  %d MHz" % (d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e, open("gpurun_out/r2_%s.err" % sys.argv[1]).read()[-300:])
PY
done
done
