mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r02m.log 2>&1; echo smoke rc=$?; tail -5 gpurun_out/smoke_r02m.log
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_drivers.py -q -m gpu --timeout 600 -rxXs -s -k "x3 or golden or config or u8" > gpurun_out/tests_x3_r02m.log 2>&1; echo tests rc=$?; grep "measured\|passed\|failed\|FAILED\|Error" gpurun_out/tests_x3_r02m.log | grep -i "x3\|passed\|failed\|error" | tail -50
