mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 600 -x -k "single_conv" > gpurun_out/tests_r02z3.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/tests_r02z3.log
