( timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x ) > gpurun_out/x_pytest.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/x_pytest.log
bash tools/gpu_ab.sh RESNET_B200_PACK_ASIDE "0 1" "c2 c4 c5" 2
