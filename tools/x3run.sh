mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -rxXs -s -k "x3 or overall_statistics or levels or losses" > gpurun_out/tests_x3_r02n.log 2>&1; echo tests rc=$?; grep "measured\|passed\|failed\|FAILED\|Error" gpurun_out/tests_x3_r02n.log | grep -i "passed\|failed\|error" | tail -50
python tools/layer_report.py --precision fp16x3 --encoder > gpurun_out/layers_enc_fp16x3_r02n.log 2>&1; tail -12 gpurun_out/layers_enc_fp16x3_r02n.log | head -4
