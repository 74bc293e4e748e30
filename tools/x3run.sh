mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -rxXs -s -k "x3 or u8 or upsample" > gpurun_out/tests_r03a.log 2>&1; echo tests rc=$?; grep "measured\|passed\|failed\|FAILED\|Error" gpurun_out/tests_r03a.log | grep -i "decoder\|config\|style_transfer sq40\|passed\|failed\|error" | tail -20
python tools/layer_report.py --precision fp16x3 > gpurun_out/layers_x3_r03a.log 2>&1; cat gpurun_out/layers_x3_r03a.log | sed -n '1,3p;11,21p'
