mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -x -k "x3 or u8" > gpurun_out/tests_r02u.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_r02u.log
python tools/layer_report.py --precision fp16x3 > gpurun_out/layers_x3_r02u.log 2>&1; cat gpurun_out/layers_x3_r02u.log | sed -n '1,3p;19,21p'
