mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -x -k "golden or u8 or totensor or reproducible or deterministic or config1" > gpurun_out/tests_r02s.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_r02s.log
python tools/layer_report.py > gpurun_out/layers_r02s.log 2>&1; head -4 gpurun_out/layers_r02s.log
python tools/layer_report.py --precision fp16x3 --encoder > gpurun_out/layers_enc_x3_r02s.log 2>&1; head -4 gpurun_out/layers_enc_x3_r02s.log
