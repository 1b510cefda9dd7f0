# compute-sanitizer over the hot path: the smoke step (miniature network, tf32 + bf16) and one batch-32 ResNet-50 step per dtype.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh memcheck'      (one tool per gpurun call: B200_PROFILING.md)
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh racecheck'
TOOL=${1:-memcheck}
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitizer_$TOOL.txt
: > $OUT
run() {
  echo "== $TOOL: $*" | tee -a $OUT
  "$@" > /dev/null 2>&1 || { echo "plain run failed: $*" | tee -a $OUT; return; }     # the same command without the tool first
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 "$@" > gpurun_out/san_tmp.log 2>&1
  echo "exit $?" | tee -a $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race reported|hazard|smoke ok|images/s|img/s" gpurun_out/san_tmp.log | head -n 30 | tee -a $OUT
}
run python -c "import __graft_entry__ as g; g.smoke()"
run python tools/one_step.py --batch 32 --dtype tf32 --steps 1
run python tools/one_step.py --batch 32 --dtype bf16 --steps 1
