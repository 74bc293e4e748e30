mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_r03i.log 2> gpurun_out/bench_8gpu_r03i.err; echo rc=$?; tail -c 600 gpurun_out/bench_8gpu_r03i.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_check.py --images 256 --size 256 --batch 16 --precision fp16x3 --oracle > gpurun_out/multi_gpu_check_8gpu_r03i.log 2>&1; echo rc=$?; tail -1 gpurun_out/multi_gpu_check_8gpu_r03i.log
