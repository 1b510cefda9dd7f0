# Round 2, second GPU call: wgrad work-item order A/B, multi-issuer validation + potential
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --dtype f32 --passes 3 --variants "RESNET_B200_WGRAD_SPLIT_MAJOR=0" > gpurun_out/r2b_wgrad_f32.txt 2>&1; echo "wgrad f32 exit $?"
timeout 300 python tools/conv_bench.py --dtype bf16 --passes 3 --variants "RESNET_B200_WGRAD_SPLIT_MAJOR=0" > gpurun_out/r2b_wgrad_bf16.txt 2>&1; echo "wgrad bf16 exit $?"
RESNET_B200_TEST_ISSUERS=1 timeout 300 python -m pytest tests/test_gpu_bf16.py -q --timeout 60 -k "issuer_counts" > gpurun_out/r2b_issuers_pytest.log 2>&1; echo "issuer tests exit $?"; tail -n 3 gpurun_out/r2b_issuers_pytest.log
timeout 400 python tools/conv_bench.py --dtype f32 --passes 0,1,3 --variants "RESNET_B200_ISSUERS=2;RESNET_B200_ISSUERS=4" > gpurun_out/r2b_issuers_f32.txt 2>&1; echo "issuers f32 exit $?"
timeout 400 python tools/conv_bench.py --dtype bf16 --passes 0,1,3 --variants "RESNET_B200_ISSUERS=2;RESNET_B200_ISSUERS=4" > gpurun_out/r2b_issuers_bf16.txt 2>&1; echo "issuers bf16 exit $?"
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r2b_pytest.log
timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench c2 exit $?"
tail -n 2 gpurun_out/r2b_wgrad_f32.txt gpurun_out/r2b_wgrad_bf16.txt gpurun_out/r2b_issuers_f32.txt gpurun_out/r2b_issuers_bf16.txt
cut -c1-200 gpurun_out/r2b_bench_c2.json
