mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 600 -x -k "reproducible" > gpurun_out/tests_r03j.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_r03j.log
