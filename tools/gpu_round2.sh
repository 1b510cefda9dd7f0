mkdir -p gpurun_out
timeout 900 python -m pytest -q -s --timeout 600 tests/test_gpu_network.py -k "full_geometry" > gpurun_out/fullgeo.log 2>&1
tail -n 30 gpurun_out/fullgeo.log
