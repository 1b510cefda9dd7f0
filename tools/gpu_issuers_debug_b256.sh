mkdir -p gpurun_out
run() { echo "== $*"; env "$@" CUDA_LAUNCH_BLOCKING=1 RESNET_B200_TRACE=1 timeout 120 python tools/one_step.py --batch 256 --dtype bf16 --steps 1 > gpurun_out/dbg.out 2> gpurun_out/dbg.err; echo "rc=$?"; tail -n 2 gpurun_out/dbg.out; grep -c "tc_run" gpurun_out/dbg.err; grep "tc_run\|error" gpurun_out/dbg.err | tail -n 4; }
run RESNET_B200_ISSUERS_K=2 RESNET_B200_ISSUERS_W=1
run RESNET_B200_ISSUERS_K=1 RESNET_B200_ISSUERS_W=2
