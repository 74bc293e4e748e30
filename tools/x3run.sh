mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_drivers.py -q -m gpu --timeout 600 -rxXs -x > gpurun_out/tests_u8_r02o.log 2>&1; echo tests rc=$?; tail -15 gpurun_out/tests_u8_r02o.log
