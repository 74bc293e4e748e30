mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -x -k "single_conv or golden or u8 or reproducible or deterministic or config1 or smallest or upsample or saturation" > gpurun_out/tests_r02z.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_r02z.log
python tools/layer_report.py > gpurun_out/layers_r02z.log 2>&1; sed -n '1p;18,21p' gpurun_out/layers_r02z.log
python tools/layer_report.py > gpurun_out/layers_r02z2.log 2>&1; sed -n '1p;18,21p' gpurun_out/layers_r02z2.log
