mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -x -k "allreduce" > gpurun_out/tests_r03h.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_r03h.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/multi_gpu_check.py --images 64 --size 128 --batch 8 --precision fp16x3 --oracle > gpurun_out/multi_gpu_check_2gpu_r03h.log 2>&1; echo rc=$?; tail -2 gpurun_out/multi_gpu_check_2gpu_r03h.log
