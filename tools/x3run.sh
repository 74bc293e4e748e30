mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -rxXs -s -k "x3 or u8" > gpurun_out/tests_x3_r02q.log 2>&1; echo tests rc=$?; grep "measured\|passed\|failed\|FAILED\|Error" gpurun_out/tests_x3_r02q.log | grep -i "decoder\|config\|u8\|passed\|failed\|error" | tail -30
python tools/layer_report.py --precision fp16x3 > gpurun_out/layers_fp16x3_r02q.log 2>&1; tail -21 gpurun_out/layers_fp16x3_r02q.log
