mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 90 python tools/one_step.py --batch 32 --dtype bf16 2>&1 | grep -v "^\[tc_run\]\|^\[k\]" | tail -n 2; echo "rc=$?"; }
run RESNET_B200_ISSUERS_K=1 RESNET_B200_ISSUERS_W=1
run RESNET_B200_ISSUERS_K=2 RESNET_B200_ISSUERS_W=1
run RESNET_B200_ISSUERS_K=1 RESNET_B200_ISSUERS_W=2
run RESNET_B200_ISSUERS_K=2 RESNET_B200_ISSUERS_W=1 RESNET_B200_EPI_GROUPS=1
run RESNET_B200_ISSUERS_K=2 RESNET_B200_ISSUERS_W=1 RESNET_B200_RESIDENT_B=0
run RESNET_B200_ISSUERS_K=2 RESNET_B200_ISSUERS_W=1 RESNET_B200_FUSED_STATS=0
timeout 400 python -m pytest tests/test_gpu_bf16.py -q --timeout 120 -k "issuer_counts" 2>&1 | tail -n 15
