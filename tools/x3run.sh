mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -x -k "single_conv or golden or reproducible or deterministic or config1 or smallest or upsample" > gpurun_out/tests_r03d.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_r03d.log
